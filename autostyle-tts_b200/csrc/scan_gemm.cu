// K3 placeholder until the tcgen05 kernel lands in this file.
#include "avs_internal.h"

int avs_launch_scan_gemm(avs_store* s, int nq, const AvsLevel& lv, int cap, cudaStream_t st) {
    (void)s; (void)nq; (void)lv; (void)cap; (void)st;
    avs_set_error("tensor-core scan is not built into this library");
    return AVS_E_STATE;
}
void avs_gemm_state_free(avs_store* s) { s->gemm_state = nullptr; }
