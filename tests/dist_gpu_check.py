"""Multi-GPU parity check for the row-sharded search (run under torchrun on the GPU box; one
process per GPU, NCCL).  Not collected by pytest: the `-m gpu` suite runs on one GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py

Every rank fills its contiguous shard of the synthetic stream on device, all ranks search the same
query batch (local exact top-k -> one ncclAllGather -> merge), and the merged result must equal the
oracle's answer on the whole database, on every rank, for both scan paths.
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import flat_search as fs  # noqa: E402

sharded = importlib.import_module("autostyle-tts_b200.sharded")
synth = importlib.import_module("autostyle-tts_b200.synth")


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for (n, d, nq, k, metric) in [(200_003, 256, 5, 10, "COSINE"), (150_000, 768, 96, 10, "COSINE"),
                                  (64_000, 128, 300, 100, "IP"), (1000, 64, 3, 50, "COSINE"),
                                  (250_000, 128, 1, 10, "COSINE"), (250_000, 128, 2, 100, "COSINE")]:   # gemv, dense 64 K-row level
        ss = sharded.ShardedStore(d, metric, n, rank, world, device=local)
        ss.fill_synthetic(42)
        Q = synth.planted_queries(43, 42, n, nq, d)
        qd = torch.from_numpy(Q).cuda()
        X = np.concatenate([synth.synth_rows(42, lo, min(50_000, n - lo), d) for lo in range(0, n, 50_000)])
        exp_ids, exp_d, _ = fs.search_large(X, np.arange(n), Q, k, metric)
        flag = torch.tensor([1], device="cuda")
        modes = (("p2p", 1), ("nccl", 0)) if ss.p2p else (("nccl", 0),)
        for name, val in modes:                       # fused peer-memory exchange+merge, then ncclAllGather + merge
            ss.store.set_option("p2p_merge", val)
            for rep in range(3):                      # repeated searches exercise the alternating exchange buffers
                ids, sc = ss.search(qd, k)
            ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
            good = np.array_equal(ids, exp_ids) and np.all(np.abs(sc - exp_d) <= 1e-5 * np.maximum(1, np.abs(exp_d)))
            if not good:
                flag.zero_()
        # end to end from host buffers: 1/world of the batch copied per rank, NVLink all-gather of the slices, search, merge
        ss.store.set_option("p2p_merge", 1 if ss.p2p else 0)
        for rep in range(3):
            h_ids, h_sc = ss.search(Q, k)
        good = np.array_equal(h_ids, exp_ids) and np.all(np.abs(h_sc - exp_d) <= 1e-5 * np.maximum(1, np.abs(exp_d)))
        if not good:
            flag.zero_()
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"world={world} n={n} d={d} nq={nq} k={k} {metric}: {'OK' if flag.item() else 'MISMATCH'} "
                  f"(modes {[m[0] for m in modes] + ['host']}, shard rows {len(ss.store)}, path {ss.store.stat('last_scan_path')}, "
                  f"uncertified {ss.store.stat('uncertified_queries')}, p2p timeouts {ss.store.stat('p2p_timeouts')})", flush=True)
        ok &= bool(flag.item())
        ss.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
