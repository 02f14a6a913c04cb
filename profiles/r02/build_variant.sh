#!/bin/bash
# Builds libavs.so of an earlier commit for A/B runs on the same GPU box:  build_variant.sh <commit> <tag>
# -> profiles/r02/variants/libavs_<tag>.so (git-ignored, travels with gpurun); run with AVS_LIB=<that path>.
set -e
C=$1; T=$2; D=/tmp/avs_variant_$T
rm -rf $D; mkdir -p $D/autostyle-tts_b200/csrc $D/include
for f in $(git ls-tree --name-only $C autostyle-tts_b200/csrc/); do git show $C:$f > $D/$f; done
git show $C:include/avs.h > $D/include/avs.h
mkdir -p profiles/r02/variants
OBJS=""
for src in $D/autostyle-tts_b200/csrc/*.cu; do
  o=${src%.cu}.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I/usr/include -c $src -o $o &
  OBJS="$OBJS $o"
done
wait
/usr/local/cuda/bin/nvcc -shared -o profiles/r02/variants/libavs_$T.so $OBJS -gencode arch=compute_100a,code=sm_100a -lcudart -ldl
ls -la profiles/r02/variants/libavs_$T.so
