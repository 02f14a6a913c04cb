// Search pipeline around the scan kernels:
//   prep_queries -> [scan level l -> select level l]* -> finalize (float64 rescoring of K' + certificate)
//   -> wide_rescore -> repair_scan / repair_finalize (device-gated: exit at once unless a query was flagged)
// Which kernel scans a level (avs_search_local): the tensor-core scan (K3) for every level of a batch of 9+ queries;
// for <= 8 queries the hybrid pipeline - threshold-free level on the warp-dot kernel (K2, up to 64 K rows kept densely),
// later levels on K3's single-CTA variant; K2 alone for 1-2 queries of D > 1024 and for stores of <= 64 K rows.
// Replaces the arithmetic behind `MilvusClient.search`
// (/root/reference/milvus/search_embeddings.py:15-22, /root/reference/milvus/RAG.py:383-390).
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "avs_internal.h"

#define ST_OVERFLOW 1
#define ST_CERT_FAIL 4
#define ST_UNCERTIFIED 8

static int pow2ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ---------------------------------------------------------------------------------------------
// Query preparation.  One CTA per (padded) query slot.  COSINE: qf = q / ||q|| (norm in float64).
// Writes the fp32 copy for the gemv scan, its bf16 rounding for the tensor-core scan, ||q|| in
// float64 for the exact rescoring and the two rigorous certificate slacks:
//   |scan score - exact score| <= ||q_scan|| * r_max + ||q_scan - qf|| * xmax + accumulation bound
// with r_max = max_j ||bf16(x^_j) - x^_j|| (maintained by K1) and xmax = max ||x^_j||.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_f64(double v, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    return t;
}

__global__ void __launch_bounds__(128) prep_queries_kernel(const float* __restrict__ q, int nq, int dim, int dpad,
                                                           int metric, const float* __restrict__ gstat,
                                                           float* __restrict__ qf, __nv_bfloat16* __restrict__ qb,
                                                           double* __restrict__ qnorm, float* __restrict__ eps_gemv,
                                                           float* __restrict__ eps_gemm, u64* __restrict__ tau,
                                                           int* __restrict__ cnt, int* __restrict__ status,
                                                           int* __restrict__ flagged, int* __restrict__ flagged2,
                                                           unsigned int* __restrict__ gbar) {
    __shared__ double sh[4];
    avs_pdl_trigger();                         // the scan kernel may be staged and run its set-up while the queries are prepared
    const int qi = blockIdx.x;
    if (qi == 0 && threadIdx.x == 0) { flagged[0] = 0; flagged2[0] = 0; }   // repair queues of this search start empty
    if (qi == 0 && threadIdx.x < 4) gbar[threadIdx.x] = 0u;                 // grid-barrier counters of the persistent kernels
    float* of = qf + (size_t)qi * dpad;
    __nv_bfloat16* ob = qb + (size_t)qi * dpad;
    if (qi >= nq) {  // padding slot: zero vector, never accepts
        for (int c = threadIdx.x; c < dpad; c += blockDim.x) {
            of[c] = 0.f;
            ob[c] = __float2bfloat16_rn(0.f);
        }
        if (threadIdx.x == 0) { tau[qi] = ~0ull; cnt[qi] = 0; }
        return;
    }
    const float* x = q + (size_t)qi * dim;
    double ss = 0.0;
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
        double v = (double)x[c];
        ss += v * v;
    }
    ss = block_sum_f64(ss, sh);
    const double qn = sqrt(ss);
    const float scale = (metric == AVS_METRIC_COSINE) ? (qn > 0.0 ? (float)(1.0 / qn) : 0.f) : 1.0f;
    double nf = 0.0, nb = 0.0, ne = 0.0;
    for (int c = threadIdx.x; c < dpad; c += blockDim.x) {
        float v = c < dim ? x[c] * scale : 0.f;
        __nv_bfloat16 b = __float2bfloat16_rn(v);
        float bv = __bfloat162float(b);
        of[c] = v;
        ob[c] = b;
        nf += (double)v * v;
        nb += (double)bv * bv;
        ne += (double)(bv - v) * (double)(bv - v);
    }
    nf = block_sum_f64(nf, sh);
    nb = block_sum_f64(nb, sh);
    ne = block_sum_f64(ne, sh);
    if (threadIdx.x == 0) {
        const double r_max = (double)gstat[0] * 1.0001 + 1e-7;
        const double xmax = (metric == AVS_METRIC_COSINE) ? 1.00001 : (double)gstat[1] * 1.00001;
        const double n_f = sqrt(nf), n_b = sqrt(nb), e_q = sqrt(ne);
        // fp32 FMA chains (gemv) and tensor-core fp32 accumulation (gemm): n*u*|q||x| style bounds,
        // doubled; plus the fp32 rounding of the normalised operands themselves.
        const double acc_v = (((double)dpad / 32.0 + 8.0) * 1.2e-7 + 1e-6) * n_f * xmax;
        const double acc_m = ((double)dpad * 1.2e-7 + 1e-6) * n_b * xmax;
        qnorm[qi] = qn;
        const float e_v = (float)(n_f * r_max + acc_v), e_m = (float)(n_b * r_max + e_q * xmax + acc_m);
        eps_gemv[qi] = e_v;
        eps_gemm[qi] = fmaxf(e_m, e_v);   // the hybrid small-batch pipeline mixes both scans under this one slack
        tau[qi] = 0ull;
        cnt[qi] = 0;
        status[qi] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Level select (K4).  One CTA per query: bitonic sort (descending) of the collected keys in
// shared memory.  Intermediate level: the rank-j key becomes the next level's threshold and the
// survivors stay in the buffer.  Final level: the best K' keys are the candidate list and
// `bound` is an upper bound on the scan score of every row NOT in the list.
// ---------------------------------------------------------------------------------------------
__device__ void bitonic_sort_desc(u64* sm, int P) {
    for (int k2 = 2; k2 <= P; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool desc = (i & k2) == 0;
                    const u64 a = sm[i], b = sm[ixj];
                    if (desc ? (a < b) : (a > b)) { sm[i] = b; sm[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// Warp-level bitonic sort (descending) of 32*EPL keys held in registers: element e lives in lane e % 32,
// register e / 32.  Strides >= 32 are register-to-register, smaller ones one shuffle.
template <int EPL>
__device__ __forceinline__ void warp_sort_desc(u64 (&k)[EPL], int lane) {
#pragma unroll
    for (int k2 = 2; k2 <= 32 * EPL; k2 <<= 1) {
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int jr = j >> 5;
#pragma unroll
                for (int r = 0; r < EPL; ++r) {
                    const int pr = r ^ jr;
                    if (pr > r) {
                        const bool desc = (((32 * r) & k2) == 0);   // k2 >= 64 here: the bit comes from r alone
                        const u64 x = k[r], y = k[pr];
                        const bool sw = desc ? (x < y) : (x > y);
                        k[r] = sw ? y : x;
                        k[pr] = sw ? x : y;
                    }
                }
            } else {
                const bool lower = (lane & j) == 0;
#pragma unroll
                for (int r = 0; r < EPL; ++r) {
                    const bool desc = (((lane + 32 * r) & k2) == 0);
                    const u64 other = __shfl_xor_sync(0xffffffffu, k[r], j);
                    const u64 mx = k[r] > other ? k[r] : other, mn = k[r] > other ? other : k[r];
                    k[r] = (desc == lower) ? mx : mn;
                }
            }
        }
    }
}

struct SelectArgs {
    u64* cand; int* cnt; int cap; u64* tau; int j_rank; int is_final; int kprime; int64_t n_rows;
    u64* topkeys; int* topn; float* bound; int* status; int dense_total; int nq;
    const u64* dense_src; int dense_stride;   // gemv path: the threshold-free level's keys live in their own buffer
    const float* eps; int k_eps;   // k_eps > 0 on the level that sets the LAST threshold: keep it 2.5 eps under the k-th score
};

// Shared tail of both select paths: `at(e)` returns the e-th best key (0 beyond n), `store(e, key)`
// is only used by the block path.  Runs on one thread.
// `lost`: more keys were offered than the buffer they were collected in can hold (never the case for the dense buffer).
__device__ __forceinline__ void select_emit_scalar(const SelectArgs& a, int q, int total, int n, u64 key_j, u64 key_kp, bool lost) {
    if (!a.is_final) {
        int jj = a.j_rank;
        if (lost) { jj = (int)(((long long)a.j_rank * a.cap) / total); if (jj < 1) jj = 1; }
        if (n >= jj) a.tau[q] = key_j;
        a.cnt[q] = n >= jj ? jj : n;
        if (lost) a.status[q] |= ST_OVERFLOW;                // rows were lost for good: force the exact repair
    } else {
        const int m = n < a.kprime ? n : a.kprime;
        a.topn[q] = m;
        float b;
        int st = 0;
        if (lost || (a.status[q] & ST_OVERFLOW)) { b = INFINITY; st = ST_OVERFLOW; }   // lost entries
        else if (n > a.kprime) b = avs_key_score(key_kp);                 // rows outside the K' candidates <= K'-th key
        else if ((int64_t)n >= a.n_rows) b = -INFINITY;                   // every row is a candidate
        else b = a.tau[q] == 0ull ? -INFINITY : avs_key_score(a.tau[q]);  // rows outside < threshold
        a.bound[q] = b;
        a.status[q] |= st;
    }
}

// One CTA per query: block-wide bitonic sort of the collected keys in shared memory.
__global__ void __launch_bounds__(1024) select_level_kernel(SelectArgs a) {
    extern __shared__ u64 sm[];
    __shared__ int s_nz;
    const int lane = threadIdx.x & 31;
    const int qq = blockIdx.x;
    int total = a.dense_total > 0 ? a.dense_total : a.cnt[qq];
    int n = (total < a.cap || a.dense_src) ? total : a.cap;
    int P = 32;
    while (P < n) P <<= 1;
    u64* c = a.cand + (size_t)qq * a.cap;                                        // survivors are written here
    const u64* cin = a.dense_src ? a.dense_src + (size_t)qq * a.dense_stride : c;   // keys are read from here
    // Intermediate level with many keys and a small rank j (the dense sparsest level: 1024-2048 keys, j ~ 16):
    // no full sort.  The j-th largest of 64 strided group maxima is a lower bound P of the j-th largest key, the
    // keys >= P (j..a few dozen) are compacted and sorted by one warp in registers.
    if (!a.is_final && (total <= a.cap || a.dense_src) && a.j_rank <= 32 && n >= 32 * a.j_rank) {
        __shared__ u64 gmax64[64];
        __shared__ u64 lst[256];
        __shared__ int s_m;
        __shared__ u64 s_P;
        const int nt = blockDim.x, tid = threadIdx.x;
        u64 lm = 0;
        for (int i = tid; i < n; i += nt) { const u64 key = cin[i]; lm = key > lm ? key : lm; }
        sm[tid] = lm;
        if (tid == 0) s_m = 0;
        __syncthreads();
        if (tid < 64) {
            u64 g = 0;
            for (int t2 = tid; t2 < nt; t2 += 64) g = sm[t2] > g ? sm[t2] : g;
            gmax64[tid] = g;
        }
        __syncthreads();
        if (tid < 32) {
            u64 k2[2] = {gmax64[lane], gmax64[lane + 32]};
            warp_sort_desc<2>(k2, lane);
            const int e = a.j_rank - 1;
            const u64 p0 = __shfl_sync(0xffffffffu, k2[0], e & 31), p1 = __shfl_sync(0xffffffffu, k2[1], e & 31);
            if (lane == 0) s_P = e < 32 ? p0 : p1;
        }
        __syncthreads();
        const u64 P = s_P;
        if (P != 0ull) {
            for (int i = tid; i < n; i += nt) {
                const u64 key = cin[i];
                if (key >= P) { const int pos = atomicAdd(&s_m, 1); if (pos < 256) lst[pos] = key; }
            }
        }
        __syncthreads();
        const int m = s_m;
        if (P != 0ull && m <= 256) {            // m >= j_rank by construction
            if (tid < 32) {
                u64 k8[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) k8[r] = (lane + 32 * r) < m ? lst[lane + 32 * r] : 0ull;
                warp_sort_desc<8>(k8, lane);
                const int e = a.j_rank - 1;     // < 32: register 0
                const u64 tj = __shfl_sync(0xffffffffu, k8[0], e);
                if (lane < a.j_rank) c[lane] = k8[0];
                if (lane == 0) { a.tau[qq] = tj; a.cnt[qq] = a.j_rank; }
            }
            return;
        }
        __syncthreads();                        // too many keys at the pivot (ties): full sort below
    }
    // Dense buffer of the gemv path (up to 64 K keys, one CTA of 1024 threads per query) with a rank beyond the warp
    // pivot path above: block pivot.  The rank-`want` value of the 1024 thread-local maxima is a lower bound of the
    // rank-`want` key, so only the few keys above it (about `want` of them) are compacted into shared memory and
    // sorted: two passes over the keys instead of the radix select's nine, which matters when a single query waits.
    if (a.dense_src && blockDim.x == 1024 && n >= 4096 && (a.is_final ? a.kprime : a.j_rank) <= 512) {
        __shared__ int s_bm, s_bnz;
        const int nt = blockDim.x, tid = threadIdx.x;
        const int want = a.is_final ? a.kprime : a.j_rank;
        u64 lm = 0;
        int nzc = 0;
        for (int i = tid; i < n; i += nt) { const u64 key = cin[i]; lm = key > lm ? key : lm; nzc += key != 0ull; }
        sm[tid] = lm;
        if (tid == 0) { s_bm = 0; s_bnz = 0; }
        __syncthreads();
#pragma unroll
        for (int o = 16; o; o >>= 1) nzc += __shfl_xor_sync(0xffffffffu, nzc, o);
        if (lane == 0 && nzc) atomicAdd(&s_bnz, nzc);
        bitonic_sort_desc(sm, 1024);                 // ends with a barrier: s_bnz is complete as well
        const u64 Pv = sm[want - 1];
        const int n_real = s_bnz;
        __syncthreads();                             // everyone holds the pivot before the list overwrites sm
        // final level: the wide-rescoring stage wants every real key in the candidate buffer when they fit -> radix path
        const bool usable = Pv != 0ull && !(a.is_final && n_real <= a.cap);
        if (usable) {
            for (int i = tid; i < n; i += nt) {
                const u64 key = cin[i];
                if (key >= Pv) { const int pos = atomicAdd(&s_bm, 1); if (pos < a.cap) sm[pos] = key; }
            }
        }
        __syncthreads();
        const int m = s_bm;                          // >= want: the `want` largest local maxima are distinct keys
        if (usable && m <= a.cap) {
            int P2 = 32;
            while (P2 < m) P2 <<= 1;
            for (int i = m + tid; i < P2; i += nt) sm[i] = 0ull;
            __syncthreads();
            bitonic_sort_desc(sm, P2);
            if (!a.is_final) {
                for (int i = tid; i < want; i += nt) c[i] = sm[i];
                if (tid == 0) { a.tau[qq] = sm[want - 1]; a.cnt[qq] = want; }
            } else {
                for (int i = tid; i < a.kprime; i += nt) a.topkeys[(size_t)qq * a.kprime + i] = sm[i];
                if (tid == 0) {
                    a.cnt[qq] = a.cap + 1;           // the collected set does not fit the candidate buffer: no wide stage
                    select_emit_scalar(a, qq, n_real, n_real, sm[want - 1], sm[want - 1], false);
                }
            }
            return;
        }
        __syncthreads();                             // few real keys (row filter) or too many above the pivot: radix select
    }
    // Large collected sets (big K', many-row dense level): 8-pass byte-wise radix select of the rank-`want` key
    // straight from L2 (O(n) per pass, no 64-128 KB of shared memory) and an unordered compaction of the keys
    // above it.  Neither the next level nor the rescoring stage needs the survivors sorted.
    if (n > 1024 || a.dense_src) {
        __shared__ int hist[256];
        __shared__ int s_sel[2];
        __shared__ u64 lst[256];
        __shared__ int s_m;
        const int nt = blockDim.x, tid = threadIdx.x;
        int n_real = n;
        if (a.dense_total > 0) {                    // padding slots of the dense level hold key 0
            if (tid == 0) s_m = 0;
            __syncthreads();
            int nzc = 0;
            for (int i = tid; i < n; i += nt) nzc += cin[i] != 0ull;
#pragma unroll
            for (int o = 16; o; o >>= 1) nzc += __shfl_xor_sync(0xffffffffu, nzc, o);
            if (lane == 0 && nzc) atomicAdd(&s_m, nzc);
            __syncthreads();
            n_real = s_m;
            total = n_real;
            __syncthreads();
        }
        const bool lost = !a.dense_src && total > a.cap;   // the dense buffer holds every key of its level
        int jj = a.j_rank;
        if (lost) { jj = (int)(((long long)a.j_rank * a.cap) / total); if (jj < 1) jj = 1; }
        const int want = a.is_final ? (n_real < a.kprime ? n_real : a.kprime) : (n_real >= jj ? jj : n_real);
        u64 prefix = 0, maskb = 0;
        int rank = want - 1;
        if (want > 0) {
            for (int byte = 7; byte >= 0; --byte) {
                for (int i = tid; i < 256; i += nt) hist[i] = 0;
                __syncthreads();
                for (int i0 = 0; i0 < n; i0 += nt) {      // whole warps stay converged for the match below
                    const int i = i0 + tid;
                    const u64 key = i < n ? cin[i] : 0ull;
                    const bool in = i < n && (key & maskb) == prefix;
                    const int bin = in ? (int)((key >> (8 * byte)) & 0xFFull) : 256 + lane;   // non-members: unique dummies
                    // scores share their leading bytes: aggregate equal bins inside the warp, one atomic per bin
                    const unsigned peers = __match_any_sync(0xffffffffu, bin);
                    if (in && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
                }
                __syncthreads();
                if (tid < 32) {                     // lane l owns bins 255-8l .. 248-8l (descending key order)
                    int loc[8], sum = 0;
#pragma unroll
                    for (int b = 0; b < 8; ++b) { loc[b] = hist[255 - 8 * lane - b]; sum += loc[b]; }
                    int incl = sum;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int t2 = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t2; }
                    int accb = incl - sum;
                    if (rank >= accb && rank < incl) {
#pragma unroll
                        for (int b = 0; b < 8; ++b) {
                            if (rank >= accb && rank < accb + loc[b]) { s_sel[0] = 255 - 8 * lane - b; s_sel[1] = accb; }
                            accb += loc[b];
                        }
                    }
                }
                __syncthreads();
                prefix |= (u64)s_sel[0] << (8 * byte);
                maskb |= 0xFFull << (8 * byte);
                rank -= s_sel[1];
            }
        }
        const u64 Pk = want > 0 ? prefix : ~0ull;   // the rank-(want-1) key; exactly `want` keys are >= it
        if (tid == 0) s_m = 0;
        if (a.is_final) for (int i = tid; i < a.kprime; i += nt) a.topkeys[(size_t)qq * a.kprime + i] = 0ull;
        __syncthreads();
        for (int i = tid; i < n; i += nt) {
            const u64 key = cin[i];
            if (key >= Pk && key != 0ull) {
                const int pos = atomicAdd(&s_m, 1);
                if (a.is_final) { if (pos < a.kprime) a.topkeys[(size_t)qq * a.kprime + pos] = key; }
                else if (pos < 256) lst[pos] = key;
            }
        }
        __syncthreads();
        if (!a.is_final) {
            for (int i = tid; i < want && i < 256; i += nt) c[i] = lst[i];
        } else if (a.dense_src) {
            // single-level search on the gemv path: the keys live in the dense buffer; the wide-rescoring stage reads
            // the candidate buffer, so hand it every real key when they fit (else it defers to the exact scan)
            if (tid == 0) s_m = 0;
            __syncthreads();
            if (n_real <= a.cap) {
                for (int i = tid; i < n; i += nt) {
                    const u64 key = cin[i];
                    if (key != 0ull) c[atomicAdd(&s_m, 1)] = key;
                }
            }
            if (tid == 0) a.cnt[qq] = n_real <= a.cap ? n_real : a.cap + 1;
        } else if (tid == 0) {
            a.cnt[qq] = a.dense_total > 0 ? n : total;   // dense level inside the buffer: slots (zero = empty), not keys
        }
        if (tid == 0) select_emit_scalar(a, qq, total, n_real, Pk, Pk, lost);
        return;
    }
    if (threadIdx.x == 0) s_nz = 0;
    __syncthreads();
    int nz = 0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const u64 key = i < n ? c[i] : 0ull;
        sm[i] = key;
        nz += key != 0ull;
    }
    if (a.dense_total > 0) {                 // dense level: padding slots hold key 0 and sort last
#pragma unroll
        for (int o = 16; o; o >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, o);
        if (lane == 0 && nz) atomicAdd(&s_nz, nz);
    }
    __syncthreads();
    if (a.dense_total > 0) { n = s_nz; total = n; }
    bitonic_sort_desc(sm, P);
    int jj = a.j_rank;
    if (total > a.cap) { jj = (int)(((long long)a.j_rank * a.cap) / total); if (jj < 1) jj = 1; }
    if (!a.is_final) {
        int keep = n >= jj ? jj : n;
        u64 tau_new = n >= jj ? sm[jj - 1] : a.tau[qq];
        // Last threshold: besides leaving ~j*ratio survivors it must sit at least 2.5 eps below the k-th scan score
        // seen so far, so that the rows it admits are enough for the wide-rescoring certificate
        // (exact k-th >= scan k-th - eps  >  threshold + eps).  Large dims (big eps relative to the score spacing)
        // get more survivors this way, small dims are unaffected.
        if (a.k_eps > 0 && n >= a.k_eps && total <= a.cap) {
            const u64 t_eps = avs_make_key(avs_key_score(sm[a.k_eps - 1]) - 2.5f * a.eps[qq], 0xFFFFFFFFu);
            if (t_eps < tau_new) {
                tau_new = t_eps;
                int lo = keep, hi = n;                    // keys are sorted: first index with key < tau_new
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (sm[mid] >= tau_new) lo = mid + 1; else hi = mid; }
                keep = lo;
            }
        }
        for (int i = threadIdx.x; i < keep; i += blockDim.x) c[i] = sm[i];
        if (threadIdx.x == 0) {
            a.tau[qq] = tau_new;
            a.cnt[qq] = keep;
            if (total > a.cap) a.status[qq] |= ST_OVERFLOW;
        }
        return;
    }
    const int m = n < a.kprime ? n : a.kprime;
    for (int i = threadIdx.x; i < a.kprime; i += blockDim.x) a.topkeys[(size_t)qq * a.kprime + i] = i < m ? sm[i] : 0ull;
    for (int i = threadIdx.x; i < n; i += blockDim.x) c[i] = sm[i];   // sorted: the wide rescoring stage reads it
    if (threadIdx.x == 0) {
        a.cnt[qq] = total;
        select_emit_scalar(a, qq, total, n, sm[jj - 1 < P ? jj - 1 : P - 1], sm[a.kprime - 1 < P ? a.kprime - 1 : P - 1], total > a.cap);
    }
}

// ---------------------------------------------------------------------------------------------
// Exact float64 score of one master row, one warp per (query, candidate).  Canonical accumulation order:
// the row is cut into quads of 4 consecutive elements, lane l owns quads l, l+32, ... and accumulates them in
// order (elements past `dim` count as zero), then an xor-butterfly (commutative, so every lane ends with the
// same bits).  The result depends only on the row's and the query's content: identical rows tie exactly and
// fall through to the id comparison.  16-byte aligned operands take the vector path (2 x 128-bit loads of the
// row and of the query in flight per lane before the first FMA); the scalar path keeps the same order and
// therefore the same bits.  Used by both rescoring and repair.
// ---------------------------------------------------------------------------------------------
#define AVS_XS_UNROLL 2
__device__ __forceinline__ double exact_score(const float* __restrict__ x, const float* __restrict__ q, int dim,
                                              double qn, int metric, int lane) {
    double dot = 0.0, xx = 0.0;
    const int nquad = (dim + 3) >> 2;
    const bool vec = ((dim & 3) == 0) && (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(q)) & 15) == 0);
    if (vec) {
        const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
        const float4* __restrict__ q4 = reinterpret_cast<const float4*>(q);
        for (int base = 0; base < nquad; base += 32 * AVS_XS_UNROLL) {
            float4 xv[AVS_XS_UNROLL], qv[AVS_XS_UNROLL];
#pragma unroll
            for (int u = 0; u < AVS_XS_UNROLL; ++u) {
                const int i = base + 32 * u + lane;
                if (i < nquad) { xv[u] = __ldg(x4 + i); qv[u] = __ldg(q4 + i); }
                else { xv[u] = make_float4(0.f, 0.f, 0.f, 0.f); qv[u] = xv[u]; }
            }
#pragma unroll
            for (int u = 0; u < AVS_XS_UNROLL; ++u) {
                const double a0 = (double)xv[u].x, a1 = (double)xv[u].y, a2 = (double)xv[u].z, a3 = (double)xv[u].w;
                dot = fma(a0, (double)qv[u].x, dot); xx = fma(a0, a0, xx);
                dot = fma(a1, (double)qv[u].y, dot); xx = fma(a1, a1, xx);
                dot = fma(a2, (double)qv[u].z, dot); xx = fma(a2, a2, xx);
                dot = fma(a3, (double)qv[u].w, dot); xx = fma(a3, a3, xx);
            }
        }
    } else {
        for (int i = lane; i < nquad; i += 32) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = 4 * i + j;
                const double a = e < dim ? (double)__ldg(x + e) : 0.0, b = e < dim ? (double)__ldg(q + e) : 0.0;
                dot = fma(a, b, dot);
                xx = fma(a, a, xx);
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        dot += __shfl_xor_sync(0xffffffffu, dot, o);
        xx += __shfl_xor_sync(0xffffffffu, xx, o);
    }
    if (metric == AVS_METRIC_COSINE) {
        const double den = sqrt(xx) * qn;
        return den > 0.0 ? dot / den : 0.0;
    }
    return dot;
}

// Same arithmetic for NR master rows at once against one query: per row the accumulation order is exactly
// exact_score()'s (identical bits), but the NR rows' loads are in flight together and the query is loaded once -
// the rescoring gather is latency-bound (ncu: 30 % of DRAM throughput with one row per warp at a time).
template <int NR>
__device__ __forceinline__ void exact_score_rows(const float* const (&x)[NR], const float* __restrict__ q, int dim, double qn,
                                                 int metric, int lane, double (&out)[NR]) {
    double dot[NR], xx[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) { dot[r] = 0.0; xx[r] = 0.0; }
    const int nquad = dim >> 2;                       // caller guarantees dim % 4 == 0 (rows and query 16-byte aligned)
    const float4* __restrict__ q4 = reinterpret_cast<const float4*>(q);
    for (int base = 0; base < nquad; base += 64) {
        float4 qv[2], xv[NR][2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = base + 32 * u + lane;
            const bool in = i < nquad;
            qv[u] = in ? __ldg(q4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < NR; ++r) xv[r][u] = in ? __ldg(reinterpret_cast<const float4*>(x[r]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const double a0 = (double)xv[r][u].x, a1 = (double)xv[r][u].y, a2 = (double)xv[r][u].z, a3 = (double)xv[r][u].w;
                dot[r] = fma(a0, (double)qv[u].x, dot[r]); xx[r] = fma(a0, a0, xx[r]);
                dot[r] = fma(a1, (double)qv[u].y, dot[r]); xx[r] = fma(a1, a1, xx[r]);
                dot[r] = fma(a2, (double)qv[u].z, dot[r]); xx[r] = fma(a2, a2, xx[r]);
                dot[r] = fma(a3, (double)qv[u].w, dot[r]); xx[r] = fma(a3, a3, xx[r]);
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            dot[r] += __shfl_xor_sync(0xffffffffu, dot[r], o);
            xx[r] += __shfl_xor_sync(0xffffffffu, xx[r], o);
        }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        if (metric == AVS_METRIC_COSINE) {
            const double den = sqrt(xx[r]) * qn;
            out[r] = den > 0.0 ? dot[r] / den : 0.0;
        } else out[r] = dot[r];
    }
}

// ---------------------------------------------------------------------------------------------
// Finalize: one CTA per query sorts the <= 256 rescored candidates by (score desc, id asc, row asc)
// and checks the exactness certificate: every row outside the candidate list has scan score
// <= bound, hence exact score <= bound + eps; if the k-th exact score is strictly larger, the
// top-k is proven exact.  Otherwise the query is queued for the exact repair scan.
// ---------------------------------------------------------------------------------------------
struct Hit { double s; int64_t id; uint32_t row; };
__device__ __forceinline__ bool hit_better(const Hit& a, const Hit& b) {
    if (a.s != b.s) return a.s > b.s;
    if (a.id != b.id) return a.id < b.id;
    return a.row < b.row;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 1024 ? 1 : 2) finalize_kernel(const float* __restrict__ master, const int64_t* __restrict__ ids,
                                                       const float* __restrict__ qraw, const double* __restrict__ qnorm,
                                                       int dim, int metric,
                                                       const u64* __restrict__ topkeys, const int* __restrict__ topn,
                                                       const float* __restrict__ bound, const float* __restrict__ eps,
                                                       int kprime, int k, int64_t n_rows, int force_repair,
                                                       int64_t* __restrict__ out_ids, float* __restrict__ out_scores,
                                                       int64_t* __restrict__ out_rows, double* __restrict__ out_s64,
                                                       int* __restrict__ status, int* __restrict__ flagged,
                                                       double* __restrict__ rep_thr, int* __restrict__ rep_cnt,
                                                       u64* __restrict__ dstat) {
    __shared__ Hit sm[AVS_MAX_KPRIME];
    __shared__ u64 s_keys[AVS_MAX_KPRIME];
    __shared__ uint32_t s_live[AVS_MAX_KPRIME];
    __shared__ double s_cut;
    __shared__ int s_nlive;
    avs_pdl_trigger();
    avs_pdl_wait();                            // the scan (and, through it, prep) has completed
    const int q = blockIdx.x, t = threadIdx.x;
    // independent loads first (the kernel is a chain of dependent DRAM round trips: every one taken off it counts)
    const u64 key_raw = t < kprime ? topkeys[(size_t)q * kprime + t] : 0ull;
    const int n = topn[q];
    const double eps_q = (double)eps[q];
    const double qn = qnorm[q];
    const int need = (int64_t)k < n_rows ? k : (int)n_rows;
    // Which candidates can still reach the top-k?  With s_k the `need`-th best SCAN score of the list, `need` rows have
    // exact score >= s_k - eps, and a candidate whose scan score is below s_k - 2 eps has exact score < s_k - eps: it
    // cannot be among the `need` best and its 3 KB master row need not be fetched (C2: ~40 % of K' = 32).  The list
    // arrives unordered from the radix selects, so every key's rank is counted (K' <= 256 broadcast reads).
    if (t < kprime) s_keys[t] = t < n ? key_raw : 0ull;
    if (t == 0) { s_cut = -INFINITY; s_nlive = 0; }
    __syncthreads();
    int my_rank = -1;
    u64 my_key = 0ull;
    if (t < kprime) {
        my_key = s_keys[t];
        if (my_key != 0ull) {
            int r = 0;
            for (int j = 0; j < kprime; ++j) r += s_keys[j] > my_key ? 1 : 0;
            my_rank = r;
            if (r == need - 1) s_cut = (double)avs_key_score(my_key) - 2.0 * eps_q;
        }
    }
    __syncthreads();
    if (my_rank >= 0 && (my_rank < need || (double)avs_key_score(my_key) >= s_cut)) {   // the live set is a prefix in rank order
        s_live[my_rank] = avs_key_row(my_key);
        atomicMax(&s_nlive, my_rank + 1);
    }
    __syncthreads();
    const int n_live = s_nlive;
    int P = 2;
    while (P < n_live) P <<= 1;                // sorted size; P <= K'
    {   // K5: exact float64 rescoring, one warp per candidate (8 warps take the live candidates in turns)
        const int lane = t & 31, warp = t >> 5;
        const float* qp = qraw + (size_t)q * dim;
        const int nwarps = (int)(blockDim.x >> 5);
        const bool vec = ((dim & 3) == 0) && ((reinterpret_cast<uintptr_t>(master) | reinterpret_cast<uintptr_t>(qp)) & 15) == 0;
        // the scoring loops consume a row in slices (2 x 128-bit loads per lane in flight), i.e. several dependent DRAM
        // round trips per row; asking the L2 for the whole rows first turns all but the first into L2 hits
        const int row_lines = (dim * 4 + 127) >> 7;
        if (THREADS <= 256 && vec && kprime >= 4 * nwarps) {
            for (int c0 = warp * 4; c0 < n_live; c0 += 4 * nwarps)
                for (int i = 0; i < 4 && c0 + i < n_live; ++i) {
                    const char* rp = reinterpret_cast<const char*>(master + (size_t)s_live[c0 + i] * dim);
                    for (int ln = lane; ln < row_lines; ln += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + ((size_t)ln << 7)));
                }
        } else {
            for (int c = warp; c < n_live; c += nwarps) {
                const char* rp = reinterpret_cast<const char*>(master + (size_t)s_live[c] * dim);
                for (int ln = lane; ln < row_lines; ln += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + ((size_t)ln << 7)));
            }
        }
        if (THREADS <= 256 && vec && kprime >= 4 * nwarps) {
            // four candidates per warp at a time: their row loads overlap, the query is loaded once for the four
            for (int c0 = warp * 4; c0 < P; c0 += 4 * nwarps) {
                if (c0 >= n_live) {
                    if (lane < 4 && c0 + lane < P) { Hit h; h.s = -INFINITY; h.id = INT64_MAX; h.row = 0xFFFFFFFFu; sm[c0 + lane] = h; }
                    continue;
                }
                uint32_t rows4[4];
                const float* xr[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    rows4[i] = s_live[c0 + i < n_live ? c0 + i : c0];   // the group's first row stands in for padding slots
                    xr[i] = master + (size_t)rows4[i] * dim;
                }
                const int64_t my_id = (lane < 4 && c0 + lane < n_live) ? ids[rows4[lane]] : INT64_MAX;   // in flight with the rows
                double sc4[4];
                exact_score_rows<4>(xr, qp, dim, qn, metric, lane, sc4);
                if (lane < 4 && c0 + lane < P) {
                    Hit h;
                    const int c = c0 + lane;
                    if (c < n_live) { h.row = rows4[lane]; h.s = sc4[lane]; h.id = my_id; }
                    else { h.s = -INFINITY; h.id = INT64_MAX; h.row = 0xFFFFFFFFu; }
                    sm[c] = h;
                }
            }
        } else {
        for (int c = warp; c < P; c += nwarps) {
            Hit h;
            if (c < n_live) {
                h.row = s_live[c];
                h.id = ids[h.row];                                        // in flight with the row
                h.s = exact_score(master + (size_t)h.row * dim, qp, dim, qn, metric, lane);
            } else { h.s = -INFINITY; h.id = INT64_MAX; h.row = 0xFFFFFFFFu; }
            if (lane == 0) sm[c] = h;
        }
        }
    }
    __syncthreads();
    if (P <= 64) {
        // short lists (K' = 32 for the reference's limits): every hit counts the hits that beat it - one pass of broadcast
        // reads instead of the network's 15-21 barrier-separated stages; (score, id, row) is a strict total order
        Hit mine;
        int r = 0;
        if (t < n_live) {                        // the padding entries behind the live ones stay where they are
            mine = sm[t];
            for (int j = 0; j < n_live; ++j) r += hit_better(sm[j], mine) ? 1 : 0;
        }
        __syncthreads();
        if (t < n_live) sm[r] = mine;
        __syncthreads();
    } else
    for (int k2 = 2; k2 <= P; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            if (t < P) {
                const int ixj = t ^ j;
                if (ixj > t) {
                    const bool desc = (t & k2) == 0;
                    const Hit a = sm[t], b = sm[ixj];
                    if (desc ? hit_better(b, a) : hit_better(a, b)) { sm[t] = b; sm[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    if (t < k) {
        const bool valid = t < n_live;
        out_ids[(size_t)q * k + t] = valid ? sm[t].id : -1;
        out_scores[(size_t)q * k + t] = valid ? (float)sm[t].s : -INFINITY;
        if (out_rows) out_rows[(size_t)q * k + t] = valid ? (int64_t)sm[t].row : -1;
        out_s64[(size_t)q * k + t] = valid ? sm[t].s : -INFINITY;
    }
    if (t == 0) {
        const float b = bound[q];
        bool ok = n >= need;
        if (ok && need > 0 && b != -INFINITY) ok = sm[need - 1].s > (double)b + eps_q;
        if (force_repair) ok = false;
        if (!ok) {                                   // stage 1 of the repair: rescore the whole collected set
            const int pos = atomicAdd(flagged, 1);
            flagged[1 + pos] = q;
            status[q] |= ST_CERT_FAIL;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Repair stage 1 (wide rescoring).  The final level left EVERY row whose scan key reached the last
// threshold in the query's buffer (a few K' rows, sorted); the K' best were not enough to prove the
// top-k, so rescore them all in float64.  Rows outside the buffer have scan score below the threshold
// (or below the WIDE_MAX-th key), which is usually far under the k-th exact score: the certificate
// passes without touching the rest of the database.  Otherwise the query moves on to the exact scan.
// ---------------------------------------------------------------------------------------------
#define AVS_WIDE_MAX 4096
__global__ void __launch_bounds__(1024) wide_rescore_kernel(const float* __restrict__ master, const int64_t* __restrict__ ids,
                                                            const float* __restrict__ qraw, const double* __restrict__ qnorm,
                                                            int dim, int metric, const u64* __restrict__ cand,
                                                            const int* __restrict__ cnt, int cap, const u64* __restrict__ tau,
                                                            const float* __restrict__ eps, int k, int64_t n_rows, int force_repair,
                                                            int64_t* __restrict__ out_ids, float* __restrict__ out_scores,
                                                            int64_t* __restrict__ out_rows, double* __restrict__ out_s64,
                                                            int* __restrict__ status, const int* __restrict__ flagged,
                                                            int* __restrict__ flagged2, double* __restrict__ rep_thr,
                                                            int* __restrict__ rep_cnt, u64* __restrict__ dstat) {
    extern __shared__ unsigned char raw[];
    Hit* sm = reinterpret_cast<Hit*>(raw);
    avs_pdl_trigger();
    avs_pdl_wait();                            // finalize has completed: the queue of flagged queries is final
    const int nf = flagged[0];
    const int f = blockIdx.x;
    if (f >= nf) return;
    const int q = flagged[1 + f];
    const int slots = cnt[q];
    // the buffer is not necessarily sorted and may hold empty (zero) slots: rescore all of it or hand the query on
    const bool lost = slots > cap || slots > AVS_WIDE_MAX || (status[q] & ST_OVERFLOW);
    const int ns = lost ? 0 : slots;
    const u64* c = cand + (size_t)q * cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    __shared__ int s_valid;
    if (threadIdx.x == 0) s_valid = 0;
    __syncthreads();
    int P = 32;
    while (P < ns) P <<= 1;
    const double qn = qnorm[q];
    const float* qp = qraw + (size_t)q * dim;
    for (int i = warp; i < P; i += nwarps) {
        Hit h;
        const u64 key = i < ns ? c[i] : 0ull;
        if (key != 0ull) {
            h.row = avs_key_row(key);
            h.s = exact_score(master + (size_t)h.row * dim, qp, dim, qn, metric, lane);
            h.id = ids[h.row];
            if (lane == 0) atomicAdd(&s_valid, 1);
        } else { h.s = -INFINITY; h.id = INT64_MAX; h.row = 0xFFFFFFFFu; }
        if (lane == 0) sm[i] = h;
    }
    __syncthreads();
    const int n = s_valid, total = s_valid;
    __syncthreads();
    for (int k2 = 2; k2 <= P; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool desc = (i & k2) == 0;
                    const Hit a = sm[i], b = sm[ixj];
                    if (desc ? hit_better(b, a) : hit_better(a, b)) { sm[i] = b; sm[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    const int need = (int64_t)k < n_rows ? k : (int)n_rows;
    // upper bound of the scan score of every row outside the rescored set
    float b;
    if ((int64_t)total >= n_rows) b = -INFINITY;
    else b = tau[q] == 0ull ? -INFINITY : avs_key_score(tau[q]);
    bool ok = !lost && n >= need;
    if (ok && need > 0 && b != -INFINITY) ok = sm[need - 1].s > (double)b + (double)eps[q];
    if (force_repair > 1) ok = false;
    if (ok) {
        for (int t = threadIdx.x; t < k; t += blockDim.x) {
            const bool valid = t < n;
            out_ids[(size_t)q * k + t] = valid ? sm[t].id : -1;
            out_scores[(size_t)q * k + t] = valid ? (float)sm[t].s : -INFINITY;
            if (out_rows) out_rows[(size_t)q * k + t] = valid ? (int64_t)sm[t].row : -1;
            out_s64[(size_t)q * k + t] = valid ? sm[t].s : -INFINITY;
        }
        if (threadIdx.x == 0) atomicAdd(dstat + 2, 1ull);
    } else if (threadIdx.x == 0) {                   // stage 2: exact scan of the whole shard
        const int pos = atomicAdd(flagged2, 1);        // every flagged query gets a slot: the repair kernel takes them in groups
        flagged2[1 + pos] = q;
        // lower bound of the true k-th score: from this stage if it rescored anything, else from finalize's
        // K' candidates (out_s64 still holds their exact top-k); -inf when neither has k rows (the repair kernel then
        // locates the k-th best score itself)
        rep_thr[pos] = (!lost && n >= need && need > 0) ? sm[need - 1].s
                                                        : (need > 0 ? out_s64[(size_t)q * k + need - 1] : -INFINITY);
        rep_cnt[pos] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Exact repair: ONE cooperative kernel (grid barriers between its phases) that settles EVERY flagged query.
//   * Flagged queries are taken in groups of up to 8: one pass over the fp32 master computes the float64 score of a
//     row against the whole group (queries staged in shared memory, the row read once), so 36 flagged queries cost
//     5 passes, not 36.
//   * collect: rows whose exact score reaches the query's lower bound `thr` of the k-th best score go to the query's
//     slice of a shared pool (slice = min(4096, pool / flagged) entries, so any number of flagged queries fits).
//   * a query WITHOUT a usable bound (fewer than k candidates were collected: thr = -inf) or whose slice overflowed
//     gets an exact one first: two histogram passes over its float-rounded scores (12 + 12 bits of the monotone
//     key) locate the k-th best score to 2^-15 relative, which becomes `thr`; then it is collected (again).
//   * finalize: a CTA per query sorts the collected rows by (score desc, id asc, row asc) and writes the top-k.
// A query ends ST_UNCERTIFIED only if more rows than its slice holds tie with the k-th score inside that last bin.
// The kernel exits at once when nothing is flagged (the common case).
// ---------------------------------------------------------------------------------------------
#define REPAIR_THREADS 512
#define REPAIR_GMAX 8
#define REPAIR_BINS 4096

struct RepairArgs {
    const float* master; const float* q; const double* qnorm; const int64_t* ids; const uint32_t* filt;
    int64_t n_rows, n_eff; int dim, metric, k, group;
    int slice_max;                                                        // largest slice per flagged query: AVS_REPAIR_CAP, or up to
                                                                          // AVS_LARGE_SLICE_MAX (a power of two) for limits > AVS_MAX_KPRIME
    const int* flagged2; double* rep_thr; int* rep_cnt; int* rep_sel;     // rep_sel: [nq][2] histogram bin / rows above it
    double* pool_s; uint32_t* pool_row; int64_t pool_items;
    unsigned int* hist;                                                   // [REPAIR_GMAX][REPAIR_BINS]
    int64_t* out_ids; float* out_scores; int64_t* out_rows; double* out_s64; int* status; u64* dstat;
    unsigned int* gbar; unsigned int* err;
};

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// grid-wide barrier of a cooperatively launched kernel (the whole grid is resident); every thread calls it
__device__ __forceinline__ void grid_barrier_all(unsigned int* gbar, unsigned int& epoch, unsigned int* err) {
    __threadfence();
    __syncthreads();
    epoch += 1;
    if (threadIdx.x == 0) {
        const unsigned int target = epoch * gridDim.x;
        atomicAdd(gbar, 1u);
        long long spins = 0;
        while (ld_acquire_gpu_u32(gbar) < target) {
            __nanosleep(64);
            if (++spins > (1ll << 24)) { atomicAdd(err, 1u); break; }
        }
        __threadfence();
    }
    __syncthreads();
}

// float64 scores of one master row against `ng` queries held in shared memory, same canonical accumulation order per
// (row, query) as exact_score(): identical bits, so hits rescored by finalize_kernel and by the repair tie exactly.
template <int GMAX>
__device__ __forceinline__ void exact_scores_group(const float* __restrict__ x, const float* __restrict__ qs, int qstride, int ng,
                                                   int dim, const double* __restrict__ qn, int metric, int lane, double (&out)[GMAX]) {
    double dot[GMAX], xx = 0.0;
#pragma unroll
    for (int g = 0; g < GMAX; ++g) dot[g] = 0.0;
    const int nquad = dim >> 2;                       // caller guarantees dim % 4 == 0 and 16-byte alignment
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
    for (int base = 0; base < nquad; base += 64) {
        float4 xv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = base + 32 * u + lane;
            xv[u] = i < nquad ? __ldg(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = base + 32 * u + lane;
            if (i < nquad) {
                const double a0 = (double)xv[u].x, a1 = (double)xv[u].y, a2 = (double)xv[u].z, a3 = (double)xv[u].w;
                xx = fma(a0, a0, xx); xx = fma(a1, a1, xx); xx = fma(a2, a2, xx); xx = fma(a3, a3, xx);
#pragma unroll
                for (int g = 0; g < GMAX; ++g) {
                    if (g < ng) {
                        const float4 qv = reinterpret_cast<const float4*>(qs + (size_t)g * qstride)[i];
                        dot[g] = fma(a0, (double)qv.x, dot[g]);
                        dot[g] = fma(a1, (double)qv.y, dot[g]);
                        dot[g] = fma(a2, (double)qv.z, dot[g]);
                        dot[g] = fma(a3, (double)qv.w, dot[g]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        xx += __shfl_xor_sync(0xffffffffu, xx, o);
#pragma unroll
        for (int g = 0; g < GMAX; ++g) dot[g] += __shfl_xor_sync(0xffffffffu, dot[g], o);
    }
    const double nx = sqrt(xx);
#pragma unroll
    for (int g = 0; g < GMAX; ++g) {
        if (metric == AVS_METRIC_COSINE) {
            const double den = nx * (g < ng ? qn[g] : 0.0);
            out[g] = den > 0.0 ? dot[g] / den : 0.0;
        } else out[g] = dot[g];
    }
}

__global__ void __launch_bounds__(REPAIR_THREADS) repair_kernel(RepairArgs a) {
    extern __shared__ unsigned char raw[];
    avs_pdl_wait();                            // wide rescoring has completed
    int nf = a.flagged2[0];
    if (nf == 0) return;                                                  // uniform: nobody reaches a barrier
    __shared__ double s_thr[REPAIR_GMAX], s_qn[REPAIR_GMAX];
    __shared__ int s_f[REPAIR_GMAX], s_q[REPAIR_GMAX], s_bin[REPAIR_GMAX];
    const int need = (int64_t)a.k < a.n_eff ? a.k : (int)a.n_eff;
    const bool large = a.slice_max > AVS_REPAIR_CAP;                      // slices beyond shared memory: sorted in place in the pool
    const int64_t nf_max = a.pool_items / (large ? a.slice_max : AVS_MAX_KPRIME);
    if (nf > nf_max) {                                                    // more flagged queries than the pool has minimal slices for
        if (blockIdx.x == 0)
            for (int f = (int)nf_max + threadIdx.x; f < nf; f += blockDim.x) { a.status[a.flagged2[1 + f]] |= ST_UNCERTIFIED; atomicAdd(a.dstat + 1, 1ull); }
        nf = (int)nf_max;
    }
    int capq = (int)(a.pool_items / nf);
    if (capq > a.slice_max) capq = a.slice_max;
    if (large) { int p2 = 1; while (p2 * 2 <= capq) p2 *= 2; capq = p2; }   // the in-place bitonic sort needs a power of two
    unsigned int epoch = 0;
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = (int64_t)blockIdx.x * (REPAIR_THREADS / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (REPAIR_THREADS / 32);
    const bool vec = (a.dim & 3) == 0;                                    // master rows and the staged queries are 16-byte aligned then
    float* qs = reinterpret_cast<float*>(raw);
    const int qstride = (a.dim + 3) & ~3;

    // mode 0: collect rows with score >= thr; mode 1 / 2: first / second histogram pass of the threshold search
    auto scan = [&](int ng, unsigned mask, int mode) {
        for (int64_t r = gwarp; r < a.n_rows; r += nwarps) {
            if (a.filt && !((a.filt[r >> 5] >> (r & 31)) & 1u)) continue;
            double sc[REPAIR_GMAX];
            const float* x = a.master + (size_t)r * a.dim;
            if (vec) exact_scores_group<REPAIR_GMAX>(x, qs, qstride, ng, a.dim, s_qn, a.metric, lane, sc);
            else {
#pragma unroll
                for (int g = 0; g < REPAIR_GMAX; ++g)
                    sc[g] = g < ng ? exact_score(x, a.q + (size_t)s_q[g] * a.dim, a.dim, s_qn[g], a.metric, lane) : 0.0;
            }
            if (lane == 0) {
#pragma unroll
                for (int g = 0; g < REPAIR_GMAX; ++g) {
                    if (g >= ng || !((mask >> g) & 1u)) continue;
                    if (mode == 0) {
                        if (sc[g] >= s_thr[g]) {
                            const int pos = atomicAdd(a.rep_cnt + s_f[g], 1);
                            if (pos < capq) {
                                a.pool_s[(size_t)s_f[g] * capq + pos] = sc[g];
                                a.pool_row[(size_t)s_f[g] * capq + pos] = (uint32_t)r;
                            }
                        }
                    } else {
                        const uint32_t key = avs_f2ord((float)sc[g]);
                        if (mode == 1) atomicAdd(a.hist + g * REPAIR_BINS + (key >> 20), 1u);
                        else if ((int)(key >> 20) == s_bin[g]) atomicAdd(a.hist + g * REPAIR_BINS + ((key >> 8) & 0xFFFu), 1u);
                    }
                }
            }
        }
    };
    // histogram of the queries in `mask` -> (bin holding the `need`-th best, rows above that bin); block 0 writes rep_sel
    auto hist_pass = [&](int ng, unsigned mask, int mode) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < REPAIR_GMAX * REPAIR_BINS; i += gridDim.x * blockDim.x) a.hist[i] = 0u;
        grid_barrier_all(a.gbar, epoch, a.err);
        scan(ng, mask, mode);
        grid_barrier_all(a.gbar, epoch, a.err);
        if (blockIdx.x == 0 && (threadIdx.x >> 5) < ng && ((mask >> (threadIdx.x >> 5)) & 1u)) {
            const int g = threadIdx.x >> 5;
            const int above0 = mode == 1 ? 0 : a.rep_sel[2 * s_f[g] + 1];
            int acc = above0, found = -1, found_above = 0;
            for (int b0 = REPAIR_BINS - 32; b0 >= 0 && found < 0; b0 -= 32) {           // bins from the top, 32 at a time
                const int v = (int)a.hist[g * REPAIR_BINS + b0 + (31 - lane)];           // lane 0 = highest bin of the chunk
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t2 = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t2; }
                const unsigned hit = __ballot_sync(0xffffffffu, acc + incl >= need);
                if (hit) {
                    const int l0 = __ffs(hit) - 1;
                    found = b0 + (31 - l0);
                    found_above = acc + __shfl_sync(0xffffffffu, incl - v, l0);
                }
                acc += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) {
                if (found < 0) { found = 0; found_above = acc; }                        // fewer than `need` rows in all: take everything
                a.rep_sel[2 * s_f[g]] = found;
                a.rep_sel[2 * s_f[g] + 1] = found_above;
            }
        }
        grid_barrier_all(a.gbar, epoch, a.err);
    };
    // exact threshold for the queries in `mask`: k-th best float-rounded score located to 2^-15 relative
    auto find_threshold = [&](int ng, unsigned mask) {
        hist_pass(ng, mask, 1);
        if (threadIdx.x < ng) s_bin[threadIdx.x] = a.rep_sel[2 * s_f[threadIdx.x]];
        __syncthreads();
        hist_pass(ng, mask, 2);
        if (threadIdx.x < ng && ((mask >> threadIdx.x) & 1u)) {
            const uint32_t key = ((uint32_t)s_bin[threadIdx.x] << 20) | ((uint32_t)a.rep_sel[2 * s_f[threadIdx.x]] << 8);
            const double edge = (double)avs_ord2f(key);
            // rows counted have float(score) >= edge, i.e. score >= edge - half an fp32 ulp: loosen by 2^-22 relative
            const double thr = edge - fabs(edge) * 2.4e-7 - 1e-300;
            s_thr[threadIdx.x] = (a.rep_sel[2 * s_f[threadIdx.x]] == 0 && s_bin[threadIdx.x] == 0) ? -INFINITY : thr;
            if (blockIdx.x == 0) { a.rep_thr[s_f[threadIdx.x]] = s_thr[threadIdx.x]; a.rep_cnt[s_f[threadIdx.x]] = 0; }
        }
        grid_barrier_all(a.gbar, epoch, a.err);
    };

    for (int g0 = 0; g0 < nf; g0 += a.group) {
        const int ng = nf - g0 < a.group ? nf - g0 : a.group;
        __syncthreads();
        if (threadIdx.x < ng) {
            const int f = g0 + threadIdx.x, qi = a.flagged2[1 + f];
            s_f[threadIdx.x] = f; s_q[threadIdx.x] = qi; s_qn[threadIdx.x] = a.qnorm[qi]; s_thr[threadIdx.x] = a.rep_thr[f];
        }
        if (vec)
            for (int i = threadIdx.x; i < ng * qstride; i += blockDim.x) {
                const int g = i / qstride, c = i - g * qstride;
                qs[i] = c < a.dim ? a.q[(size_t)a.flagged2[1 + g0 + g] * a.dim + c] : 0.f;
            }
        __syncthreads();
        unsigned need_thr = 0;                                            // no usable bound: find one first
        for (int g = 0; g < ng; ++g) if (!(s_thr[g] > -INFINITY)) need_thr |= 1u << g;
        if (need_thr) find_threshold(ng, need_thr);
        scan(ng, (1u << ng) - 1, 0);
        grid_barrier_all(a.gbar, epoch, a.err);
        unsigned over = 0;                                                // slice overflowed: the bound was too loose
        for (int g = 0; g < ng; ++g) if (a.rep_cnt[s_f[g]] > capq && !((need_thr >> g) & 1u)) over |= 1u << g;
        if (over) {
            grid_barrier_all(a.gbar, epoch, a.err);                       // everyone has read the counts before they are reset
            find_threshold(ng, over);
            scan(ng, over, 0);
            grid_barrier_all(a.gbar, epoch, a.err);
        }
    }

    // ---- finalize: one CTA per flagged query ----
    Hit* sm = reinterpret_cast<Hit*>(raw);
    for (int f = blockIdx.x; f < nf; f += gridDim.x) {
        __syncthreads();
        const int q = a.flagged2[1 + f];
        const int total = a.rep_cnt[f];
        if (total > capq || total < need) {            // more ties at the k-th score than the slice holds: keep the earlier output
            if (threadIdx.x == 0) { a.status[q] |= ST_UNCERTIFIED; atomicAdd(a.dstat + 1, 1ull); }
            continue;
        }
        int P = 32;
        while (P < total) P <<= 1;
        if (P > AVS_REPAIR_CAP) {
            // Limits beyond 2 K rows (MilvusClient allows 16 384): the slice does not fit shared memory, so the CTA sorts
            // it where it lies - same network, same order (score desc, id asc, row asc); the ids are only looked up when two
            // scores are equal.  ~120 network stages over L2-resident data: well under a millisecond per query.
            double* ps = a.pool_s + (size_t)f * capq;
            uint32_t* pr = a.pool_row + (size_t)f * capq;
            for (int i = total + threadIdx.x; i < P; i += blockDim.x) { ps[i] = -INFINITY; pr[i] = 0xFFFFFFFFu; }
            __syncthreads();
            for (int k2 = 2; k2 <= P; k2 <<= 1) {
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    for (int i = threadIdx.x; i < P; i += blockDim.x) {
                        const int ixj = i ^ j;
                        if (ixj > i) {
                            const bool desc = (i & k2) == 0;
                            const double sx = ps[i], sy = ps[ixj];
                            const uint32_t rx = pr[i], ry = pr[ixj];
                            bool x_better;                           // hit_better(x, y)
                            if (sx != sy) x_better = sx > sy;
                            else {
                                const int64_t ix = rx == 0xFFFFFFFFu ? INT64_MAX : a.ids[rx], iy = ry == 0xFFFFFFFFu ? INT64_MAX : a.ids[ry];
                                x_better = ix != iy ? ix < iy : rx < ry;
                            }
                            const bool same = sx == sy && rx == ry;  // two padding entries
                            if (!same && (desc ? !x_better : x_better)) { ps[i] = sy; pr[i] = ry; ps[ixj] = sx; pr[ixj] = rx; }
                        }
                    }
                    __syncthreads();
                }
            }
            for (int t = threadIdx.x; t < a.k; t += blockDim.x) {
                const bool valid = t < total;
                const uint32_t r = valid ? pr[t] : 0u;
                a.out_ids[(size_t)q * a.k + t] = valid ? a.ids[r] : -1;
                a.out_scores[(size_t)q * a.k + t] = valid ? (float)ps[t] : -INFINITY;
                if (a.out_rows) a.out_rows[(size_t)q * a.k + t] = valid ? (int64_t)r : -1;
                a.out_s64[(size_t)q * a.k + t] = valid ? ps[t] : -INFINITY;
            }
            if (threadIdx.x == 0) atomicAdd(a.dstat + 0, 1ull);
            continue;
        }
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            Hit h;
            if (i < total) { h.s = a.pool_s[(size_t)f * capq + i]; h.row = a.pool_row[(size_t)f * capq + i]; h.id = a.ids[h.row]; }
            else { h.s = -INFINITY; h.id = INT64_MAX; h.row = 0xFFFFFFFFu; }
            sm[i] = h;
        }
        __syncthreads();
        for (int k2 = 2; k2 <= P; k2 <<= 1) {
            for (int j = k2 >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < P; i += blockDim.x) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const bool desc = (i & k2) == 0;
                        const Hit x = sm[i], y = sm[ixj];
                        if (desc ? hit_better(y, x) : hit_better(x, y)) { sm[i] = y; sm[ixj] = x; }
                    }
                }
                __syncthreads();
            }
        }
        for (int t = threadIdx.x; t < a.k; t += blockDim.x) {
            const bool valid = t < total;
            a.out_ids[(size_t)q * a.k + t] = valid ? sm[t].id : -1;
            a.out_scores[(size_t)q * a.k + t] = valid ? (float)sm[t].s : -INFINITY;
            if (a.out_rows) a.out_rows[(size_t)q * a.k + t] = valid ? (int64_t)sm[t].row : -1;
            a.out_s64[(size_t)q * a.k + t] = valid ? sm[t].s : -INFINITY;
        }
        if (threadIdx.x == 0) atomicAdd(a.dstat + 0, 1ull);
    }
}

__global__ void fill_empty_kernel(int64_t* ids, float* scores, int64_t* rows, double* s64, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        ids[i] = -1;
        scores[i] = -INFINITY;
        if (rows) rows[i] = -1;
        if (s64) s64[i] = -INFINITY;
    }
}

// Limits above AVS_MAX_KPRIME: every query of the chunk [q0, q0 + n) goes straight to the exact repair kernel, without a
// bound (it locates the k-th best score itself with two histogram passes over the float64 scores).
__global__ void large_k_mark_kernel(int q0, int n, int* __restrict__ flagged2, double* __restrict__ rep_thr,
                                    int* __restrict__ rep_cnt, unsigned int* __restrict__ gbar) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { flagged2[0] = n; gbar[1] = 0u; }
    if (i < n) { flagged2[1 + i] = q0 + i; rep_thr[i] = -INFINITY; rep_cnt[i] = 0; }
}

// ---------------------------------------------------------------------------------------------
// scratch management
// ---------------------------------------------------------------------------------------------
template <typename T>
static int dev_alloc(T** p, size_t n) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (n == 0) n = 1;
    if (cudaMalloc((void**)p, n * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        avs_set_error("out of device memory allocating %zu bytes of search scratch", n * sizeof(T));
        return AVS_E_NOMEM;
    }
    return AVS_OK;
}

void avs_scratch_free(avs_store* s) {
    AvsScratch& c = s->sc;
    cudaFree(c.qf); cudaFree(c.qb); cudaFree(c.qnorm); cudaFree(c.eps_gemv); cudaFree(c.eps_gemm);
    cudaFree(c.cand); cudaFree(c.cnt); cudaFree(c.tau); cudaFree(c.topkeys); cudaFree(c.topn); cudaFree(c.status);
    cudaFree(c.out_s64); cudaFree(c.flagged); cudaFree(c.flagged2); cudaFree(c.dense_buf); cudaFree(c.rep_s); cudaFree(c.rep_row);
    cudaFree(c.rep_cnt); cudaFree(c.rep_thr); cudaFree(c.rep_sel); cudaFree(c.rep_hist); cudaFree(c.gbar); cudaFree(c.trace);
    cudaFree(c.gather_send); cudaFree(c.gather_recv);
    cudaFree(c.d_ids);
    if (c.h2d_q) cudaFree(c.h2d_q);
    if (c.h_out) cudaFreeHost(c.h_out);
    if (c.h_q) cudaFreeHost(c.h_q);
    c = AvsScratch();
}

// bound[] lives behind eps_gemm in one allocation to keep the struct small
static float* g_bound_of(AvsScratch& c) { return c.eps_gemm + c.nq_cap; }

int avs_scratch_reserve(avs_store* s, int nq_pad, int kprime, int cap, int k) {
    AvsScratch& c = s->sc;
    const bool grow_q = nq_pad > c.nq_cap, grow_kp = kprime > c.kprime_cap, grow_cap = cap > c.cap_cap, grow_k = k > c.k_cap;
    if (!(grow_q || grow_kp || grow_cap || grow_k)) return AVS_OK;
    AVS_CUDA(cudaDeviceSynchronize());
    const int nq2 = grow_q ? nq_pad : c.nq_cap, kp2 = grow_kp ? kprime : c.kprime_cap;
    const int cap2 = grow_cap ? cap : c.cap_cap, k2 = grow_k ? k : c.k_cap;
    if (grow_q) {
        AVS_CHECK(dev_alloc(&c.qf, (size_t)nq2 * s->dpad));
        AVS_CHECK(dev_alloc(&c.qb, (size_t)nq2 * s->dpad));
        AVS_CHECK(dev_alloc(&c.qnorm, (size_t)nq2));
        AVS_CHECK(dev_alloc(&c.eps_gemv, (size_t)nq2));
        AVS_CHECK(dev_alloc(&c.eps_gemm, (size_t)nq2 * 2));  // eps_gemm | bound
        AVS_CHECK(dev_alloc(&c.cnt, (size_t)nq2));
        AVS_CHECK(dev_alloc(&c.tau, (size_t)nq2));
        AVS_CHECK(dev_alloc(&c.topn, (size_t)nq2));
        AVS_CHECK(dev_alloc(&c.status, (size_t)nq2));
        AVS_CHECK(dev_alloc(&c.flagged, (size_t)nq2 + 1));
        AVS_CHECK(dev_alloc(&c.flagged2, (size_t)nq2 + 1));
        AVS_CHECK(dev_alloc(&c.rep_cnt, (size_t)nq2));
        AVS_CHECK(dev_alloc(&c.rep_thr, (size_t)nq2));
        AVS_CHECK(dev_alloc(&c.rep_sel, (size_t)nq2 * 2));
        // exact-repair pool, shared by the flagged queries of a search (slice = min(4096, pool / flagged) rows each):
        // room for 256 rows of every query of the largest batch, at least 1 M and at most 16 M rows
        size_t pool = (size_t)nq2 * AVS_MAX_KPRIME;
        if (pool < ((size_t)1 << 20)) pool = (size_t)1 << 20;
        if (pool > ((size_t)1 << 24)) pool = (size_t)1 << 24;
        AVS_CHECK(dev_alloc(&c.rep_s, pool));
        AVS_CHECK(dev_alloc(&c.rep_row, pool));
        c.pool_items = pool;
    }
    if (grow_q || grow_cap) AVS_CHECK(dev_alloc(&c.cand, (size_t)nq2 * cap2));
    if (grow_q || grow_kp) {
        AVS_CHECK(dev_alloc(&c.topkeys, (size_t)nq2 * kp2));
    }
    if (grow_q || grow_k) AVS_CHECK(dev_alloc(&c.out_s64, (size_t)nq2 * k2));
    if (!c.dense_buf) {
        AVS_CHECK(dev_alloc(&c.dense_buf, (size_t)(AVS_DENSE_MAX_NQ + 8) * AVS_DENSE_CAP));
        AVS_CHECK(dev_alloc(&c.rep_hist, (size_t)8 * 4096));
        AVS_CHECK(dev_alloc(&c.gbar, (size_t)4));
        AVS_CHECK(dev_alloc(&c.trace, (size_t)AVS_TRACE_SLOTS));
        AVS_CUDA(cudaMemset(c.trace, 0, AVS_TRACE_SLOTS * sizeof(u64)));
        AVS_CUDA(cudaMemset(c.gbar, 0, 4 * sizeof(unsigned int)));
    }
    c.nq_cap = nq2; c.kprime_cap = kp2; c.cap_cap = cap2; c.k_cap = k2;
    return AVS_OK;
}

// ---------------------------------------------------------------------------------------------
// scan timing hook (bench.py roofline): event pairs around the dominant (final-level) scan launch
// ---------------------------------------------------------------------------------------------
extern "C" int avs_scan_timing(avs_store* s, int enable_reset, double* mean_ms, int64_t* launches) {
    if (!s) { avs_set_error("avs_scan_timing: NULL store"); return AVS_E_INVALID; }
    AVS_CUDA(cudaSetDevice(s->device));
    if (mean_ms || launches) {
        double tot = 0.0;
        for (size_t i = 0; i < s->tev_used; ++i) {
            AVS_CUDA(cudaEventSynchronize(s->tev[2 * i + 1]));
            float ms = 0.f;
            AVS_CUDA(cudaEventElapsedTime(&ms, s->tev[2 * i], s->tev[2 * i + 1]));
            tot += ms;
        }
        if (mean_ms) *mean_ms = s->tev_used ? tot / (double)s->tev_used : 0.0;
        if (launches) *launches = (int64_t)s->tev_used;
    }
    if (enable_reset >= 0) {
        s->timing = enable_reset != 0;
        s->tev_used = 0;
        avs_p2p_timing_reset(s);
    }
    return AVS_OK;
}

static bool timing_begin(avs_store* s, cudaStream_t st, size_t* slot) {
    if (!s->timing || s->tev_used >= 8192) return false;
    if (2 * s->tev_used >= s->tev.size()) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return false;
        s->tev.push_back(a);
        s->tev.push_back(b);
    }
    *slot = s->tev_used++;
    cudaEventRecord(s->tev[2 * *slot], st);
    return true;
}
static void timing_end(avs_store* s, cudaStream_t st, size_t slot) { cudaEventRecord(s->tev[2 * slot + 1], st); }

// ---------------------------------------------------------------------------------------------
// orchestration
// ---------------------------------------------------------------------------------------------
int avs_search_local(avs_store* s, const float* q, int nq, int k, int64_t* out_ids, float* out_scores,
                     int64_t* out_rows, cudaStream_t st) {
    if (!s) { avs_set_error("avs_search: NULL store"); return AVS_E_INVALID; }
    if (nq < 0 || (nq > 0 && (!q || !out_ids || !out_scores))) { avs_set_error("avs_search: NULL query/output buffer"); return AVS_E_INVALID; }
    if (k < 1 || k > AVS_MAX_LIMIT) { avs_set_error("avs_search: limit %d outside [1, %d]", k, AVS_MAX_LIMIT); return AVS_E_INVALID; }
    if (nq == 0) return AVS_OK;
    if (nq > (1 << 20)) { avs_set_error("avs_search: more than 2^20 queries in one call"); return AVS_E_INVALID; }
    AVS_CUDA(cudaSetDevice(s->device));
    s->st_searches++;
    s->st_queries += nq;

    int kprime = s->opt_oversample > 0 ? pow2ceil(s->opt_oversample) : pow2ceil(2 * k + 8);
    if (kprime < 16) kprime = 16;
    if (kprime > AVS_MAX_KPRIME) kprime = AVS_MAX_KPRIME;
    if (kprime < k) kprime = pow2ceil(k);
    // Limits above AVS_MAX_KPRIME (MilvusClient allows 16 384; the reference never asks for more than 5) skip the bf16
    // scan altogether: the exact repair kernel serves them from the fp32 master (below).  Minimal scan scratch.
    const bool large_k = k > AVS_MAX_KPRIME;
    if (large_k) kprime = 16;
    int cap = pow2ceil(48 * kprime);
    if (cap < 1024) cap = 1024;
    if (cap > 16384) cap = 16384;
    const int nq_pad = (nq + 255) / 256 * 256;
    AVS_CHECK(avs_scratch_reserve(s, nq_pad, kprime, cap, k));
    AvsScratch& c = s->sc;
    float* bound = g_bound_of(c);

    const int64_t n_eff = s->filter ? s->filter_allowed : s->count;   // rows that may be returned
    if (s->count == 0 || n_eff == 0) {
        const int64_t n = (int64_t)nq * k;
        fill_empty_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out_ids, out_scores, out_rows, c.out_s64, n);
        s->st_launches++;
        AVS_CUDA(cudaGetLastError());
        return AVS_OK;
    }

    // Small batches in auto mode (at most 8 queries = one gemv pass): hybrid pipeline.  The threshold-free level is the
    // gemv path's (up to 64 K rows stored densely in its own buffer: one level fewer than the tensor-core schedule),
    // every later level streams through the single-CTA tensor-core scan, whose TMA bulk loads sustain more of the HBM
    // bandwidth than the warp-dot kernel does.  Schedule and level-0 kernel are the gemv path's (use_gemm = false).
    // Measured (profiles/r01/hybrid_*.json): for 1-2 queries the tensor-core scan streams D <= 1024 rows 3-14 % faster than
    // the warp-dot kernel and loses 5 % at D = 3072; for 3-8 queries the level saved (~35 us) pays off while the final
    // scan is short (C2: +10 %, 125 K-row shard: +37 %) but not on 20 GB shards, where the plain tensor-core schedule stays.
    const size_t scan_bytes = (size_t)s->count * s->dpad * 2;
    bool hybrid = s->opt_scan_path == 0 && s->opt_hybrid != 0 &&
                  ((nq <= 2 && s->dpad <= 1024) || (nq >= 3 && nq <= 8 && scan_bytes <= ((size_t)8 << 30)));
    // auto mode without the hybrid pipeline: tensor-core scan from `gemm_min_batch` queries on, except 1-2 queries of
    // D > 1024, where the warp-dot kernel streams 5 % faster
    const bool gemm_eligible = (s->opt_scan_path == 2) ||
                               (s->opt_scan_path == 0 && nq >= s->opt_gemm_min_batch && !(nq <= 2 && s->dpad > 1024));
    bool use_gemm = !hybrid && gemm_eligible;
    const bool hybrid_legacy = hybrid, use_gemm_legacy = use_gemm;
    // Sampling levels, built from the final (dense) level backwards: level l visits every stride_l-th
    // row group not visited by a sparser level; the sparsest level must fit the collection buffer with
    // threshold 0.  The tensor-core path ends with a x4 step (its epilogue pays per accepted row, so the
    // last threshold is taken from a quarter of the database); the gemv path keeps fewer, coarser levels.
    bool fine_levels = use_gemm && nq >= s->opt_fine_min_batch;   // compute-bound regime only: extra levels cost launches
    const int64_t G = (s->count + AVS_GROUP_ROWS - 1) / AVS_GROUP_ROWS;
    const int64_t rho = s->opt_ratio < 2 ? 2 : s->opt_ratio;
    int64_t strides[AVS_MAX_LEVELS];
    AvsLevel lv[AVS_MAX_LEVELS];
    int j_ranks[AVS_MAX_LEVELS];
    int L = 1;
    // Threshold ranks, from the last level backwards.  The rank-j key of what level i has collected becomes the
    // threshold of level i+1, which then ends with about j*ratio survivors (relative spread ~ 1/sqrt(j)).  The
    // last threshold must leave at least K' rows with a wide margin: K' + sigma*sqrt(K'*ratio) expected survivors;
    // every earlier level must hold comfortably more keys than the rank the next select asks for.
    auto set_ranks = [&]() {
        double need = 0.0;
        for (int i = 0; i < L; ++i) j_ranks[i] = 0;
        for (int i = L - 2; i >= 0; --i) {
            const double ratio = (double)(strides[L - 1 - i] / strides[L - 2 - i]);
            if (i == L - 2) need = (double)kprime + (double)(fine_levels ? s->opt_final_sigma : ((use_gemm || hybrid) ? s->opt_coarse_sigma : 8)) * sqrt((double)kprime * ratio);
            int64_t j = (int64_t)(need / ratio) + 1;
            if (j < 8) j = 8;
            if (j > 256) j = 256;            // the radix select hands at most 256 survivors to the next level
            j_ranks[i] = (int)j;
            need = 1.5 * (double)j + 16.0;
        }
    };
    // Sampling levels, built from the final (dense) level backwards: level l visits every stride_l-th row group not
    // visited by a sparser level.  `level0_rows`: what the threshold-free level may visit.  `fine_steps`: how many x4
    // steps sit next to the dense end on the compute-bound tensor-core schedule (its epilogue pays per accepted row,
    // so the thresholds there are refreshed often and kept tight); coarse steps for the sparser levels.
    auto build_strides = [&](int64_t level0_rows, int fine_steps, int64_t rho) {
        L = 1;
        strides[0] = 1;
        while (((G + strides[L - 1] - 1) / strides[L - 1]) * AVS_GROUP_ROWS > level0_rows && L < AVS_MAX_LEVELS) {
            int64_t r = (fine_levels && L <= fine_steps) ? s->opt_fine_ratio : rho;
            if (r == rho) {   // last coarse step: no sparser than needed to bring the threshold-free level under its row cap
                const int64_t g_prev = (G + strides[L - 1] - 1) / strides[L - 1];
                const int64_t r_needed = (g_prev * AVS_GROUP_ROWS + level0_rows - 1) / level0_rows;
                if (r_needed < r) r = r_needed < 2 ? 2 : r_needed;
            }
            strides[L] = strides[L - 1] * r;
            ++L;
        }
    };
    // Boot level (tensor-core path): the threshold-free level keeps only the AVS_BOOT_J best keys of every half group, so
    // it may visit cap / (2 J) whole groups (32 K rows at K' = 32) instead of the 2 048 rows a dense store has room for -
    // one or two levels fewer, and no warp-dot kernel + select launch in front of the scan for small batches.  Its rank
    // must not exceed J: the sparsest ratio is raised until it does not; stores too small for that keep the dense level.
    bool boot = false;
    // (a store the warp-dot path searches in ONE dense level keeps that path: nothing to save there)
    const bool boot_wanted = s->opt_boot != 0;
    if (boot_wanted && gemm_eligible && !(hybrid_legacy && s->count <= (int64_t)s->opt_dense_rows)) {
        use_gemm = true; hybrid = false;              // with a boot level the tensor-core scan takes every level, small batches too
        fine_levels = nq >= s->opt_fine_min_batch;
        // room for cap / (2 J) groups; the compute-bound schedule keeps the level to ~4 tiles per CTA pair and query block
        // sweep (its epilogue, ~10 us a tile, is slower than the MMA of the tile)
        int64_t boot_groups = cap / (2 * AVS_BOOT_J);
        if (fine_levels) {
            const int64_t n_qb = (nq + 255) / 256, pairs = s->num_sms / 2;
            const int64_t g = (4 * pairs + n_qb - 1) / n_qb;
            if (g < boot_groups) boot_groups = g < 8 ? 8 : g;
        }
        // HBM-bound batches: a level costs two grid barriers and a select, an accepted row next to nothing - allow a
        // sparser boot level (up to cap / 32: 8 * ratio expected survivors fill a quarter of the buffer at most)
        int64_t rho_boot = rho;
        if (!fine_levels) { const int64_t m = cap / 32 < 128 ? cap / 32 : 128; rho_boot = rho > m ? rho : m; }
        // Small stores on the compute-bound schedule (a shard of a strongly scaled database): when the boot level can sit
        // directly in front of the final level with a ratio of at most `boot2_ratio` and still deliver a rank <= J, the x4 level
        // between them costs more (its tiles' epilogue, two grid barriers, a select) than the looser last threshold does.
        bool two_level = false;
        if (fine_levels && s->opt_boot2_ratio >= 2 && !s->eps_rule) {
            int64_t r = (G + boot_groups - 1) / boot_groups;
            if (r < 2) r = 2;
            for (; r <= s->opt_boot2_ratio; ++r) {
                strides[0] = 1; strides[1] = r; L = 2;
                set_ranks();
                if (j_ranks[0] <= AVS_BOOT_J) { two_level = true; break; }
            }
        }
        if (!two_level) build_strides(boot_groups * AVS_GROUP_ROWS, 1, rho_boot);
        // the whole store inside the boot budget, but too large for ONE dense level: a boot level in front of the final one
        if (L == 1 && G * AVS_GROUP_ROWS > (cap < s->opt_gemm_dense_rows ? cap : s->opt_gemm_dense_rows)) {
            strides[1] = fine_levels ? s->opt_fine_ratio : 2;
            L = 2;
        }
        // the eps rule asks the select in front of the final level for rank k > J: that must not be the boot level's
        if (fine_levels && L == 2 && s->eps_rule) { strides[2] = strides[1] * 2; L = 3; }
        if (L >= 2) {
            for (int it = 0; it < 64; ++it) {
                set_ranks();
                if (j_ranks[0] <= AVS_BOOT_J) break;
                const int64_t r = strides[L - 1] / strides[L - 2];
                strides[L - 1] = strides[L - 2] * (r + (r + 3) / 4);
            }
            const int64_t ratio0 = strides[L - 1] / strides[L - 2];
            const int64_t g0 = (G + strides[L - 1] - 1) / strides[L - 1];
            boot = j_ranks[0] <= AVS_BOOT_J && g0 >= 1 && g0 <= boot_groups && (int64_t)j_ranks[0] * ratio0 * 4 <= cap;
        }
    }
    if (!boot) { use_gemm = use_gemm_legacy; hybrid = hybrid_legacy; fine_levels = use_gemm && nq >= s->opt_fine_min_batch; }
    const bool dense_gemv = !use_gemm && nq <= AVS_DENSE_MAX_NQ;
    if (!boot) {
        // the threshold-free level: <= 2048 rows inside the candidate buffer on the tensor-core path; the gemv path (few
        // queries) stores up to 64 K rows densely in its own buffer, which saves it a whole intermediate level
        const int64_t level0_rows = dense_gemv ? s->opt_dense_rows : (cap < s->opt_gemm_dense_rows ? cap : s->opt_gemm_dense_rows);
        build_strides(level0_rows, 3, rho);
        set_ranks();
    }
    for (int i = 0; i < L; ++i) {            // level 0 = sparsest
        const int64_t stride = strides[L - 1 - i];
        lv[i].stride = stride;
        lv[i].n_iter = (G + stride - 1) / stride;
        lv[i].skip = i == 0 ? 0 : strides[L - i];
        lv[i].ratio = i == 0 ? 0 : lv[i].skip / stride;
        lv[i].n_visit = lv[i].ratio > 1 ? lv[i].n_iter - (lv[i].n_iter + lv[i].ratio - 1) / lv[i].ratio : lv[i].n_iter;
        lv[i].dense = (i == 0 && boot) ? 2 :
                      (i == 0 && ((use_gemm && lv[i].n_iter * AVS_GROUP_ROWS <= cap) ||
                                  (dense_gemv && lv[i].n_iter * AVS_GROUP_ROWS <= AVS_DENSE_CAP))) ? 1 : 0;
    }

    s->st_last_final_rows = L > 1 ? (G - (G + strides[1] - 1) / strides[1]) * AVS_GROUP_ROWS : s->count;   // warp-dot path: its final level
    s->st_last_kprime = kprime;
    s->st_last_levels = L;
    s->st_last_boot = boot ? 1 : 0;
    const bool hybrid_on = hybrid && L > 1 && lv[0].dense;   // single-level searches stay on the gemv kernels
    const bool any_gemm = use_gemm || hybrid_on;
    const float* eps_used = any_gemm ? c.eps_gemm : c.eps_gemv;   // prep makes eps_gemm >= eps_gemv
    s->st_last_path = any_gemm ? 2 : 1;

    // Adaptive policy, no host synchronisation: the counters of earlier searches arrive in pinned memory whenever
    // their copy executes.  Once a query needed the exact repair scan (large dims: eps is big against the score
    // spacing), the last threshold is kept 2.5 eps under the k-th score so that wide rescoring suffices.
    if (s->h_stats && s->h_stats[0] > s->seen_repaired) { s->seen_repaired = s->h_stats[0]; s->eps_rule = true; }
    const int n_slots = any_gemm ? nq_pad : (nq + 7) / 8 * 8;   // padding slots the scan will touch
    // Kernels that synchronise their whole grid (persistent scan, repair) must not share the device with another
    // such kernel from a different stream (each could hold SMs the other waits for): chain them through one event.
    static cudaEvent_t persist_ev[64] = {};
    cudaEvent_t& pev = persist_ev[s->device & 63];
    if (!pev) AVS_CUDA(cudaEventCreateWithFlags(&pev, cudaEventDisableTiming));
    else AVS_CUDA(cudaStreamWaitEvent(st, pev, 0));
    prep_queries_kernel<<<n_slots, 128, 0, st>>>(q, nq, s->dim, s->dpad, s->metric, s->gstat, c.qf, c.qb, c.qnorm,
                                                c.eps_gemv, c.eps_gemm, c.tau, c.cnt, c.status, c.flagged, c.flagged2, c.gbar);
    s->st_launches++;
    AVS_CUDA(cudaGetLastError());

    // exact-repair kernel: queries per pass over the master (staged in shared memory) and its co-resident grid
    const int qstride = (s->dim + 3) & ~3;
    int rep_group = (96 * 1024) / (qstride * 4);
    rep_group = rep_group > REPAIR_GMAX ? REPAIR_GMAX : (rep_group < 1 ? 1 : rep_group);
    size_t rep_smem = (size_t)rep_group * qstride * 4;
    if (rep_smem < AVS_REPAIR_CAP * sizeof(Hit)) rep_smem = AVS_REPAIR_CAP * sizeof(Hit);
    static bool attr_done[64] = {};   // function attributes are per device
    static int rep_ctas_per_sm[64] = {};
    if (!attr_done[s->device & 63]) {
        AVS_CUDA(cudaFuncSetAttribute(select_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
        AVS_CUDA(cudaFuncSetAttribute(repair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        AVS_CUDA(cudaFuncSetAttribute(wide_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(AVS_WIDE_MAX * sizeof(Hit))));
        attr_done[s->device & 63] = true;
    }
    if (rep_ctas_per_sm[s->device & 63] == 0 || s->rep_smem_seen != rep_smem) {
        int per_sm = 0;
        AVS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, repair_kernel, REPAIR_THREADS, rep_smem));
        if (per_sm < 1) { avs_set_error("repair kernel does not fit an SM with %zu bytes of shared memory", rep_smem); return AVS_E_CUDA; }
        rep_ctas_per_sm[s->device & 63] = per_sm;
        s->rep_smem_seen = rep_smem;
    }

    if (large_k) {
        // slice per query: room for 2 k rows (ties inside the last histogram bin of the threshold search), a power of two;
        // the pool serves pool / slice queries per launch (at least 32), the kernel takes them 8 per pass over the master
        int slice = pow2ceil(2 * k);
        if (slice < AVS_REPAIR_CAP) slice = AVS_REPAIR_CAP;
        int chunk = (int)(c.pool_items / (size_t)slice);
        if (chunk < 1) { avs_set_error("avs_search: repair pool too small for limit %d", k); return AVS_E_STATE; }
        const int64_t n_out = (int64_t)nq * k;
        fill_empty_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(out_ids, out_scores, out_rows, c.out_s64, n_out);
        s->st_launches++;
        for (int q0 = 0; q0 < nq; q0 += chunk) {
            const int n = nq - q0 < chunk ? nq - q0 : chunk;
            large_k_mark_kernel<<<(n + 255) / 256, 256, 0, st>>>(q0, n, c.flagged2, c.rep_thr, c.rep_cnt, c.gbar);
            RepairArgs ra;
            ra.master = s->master; ra.q = q; ra.qnorm = c.qnorm; ra.ids = s->ids; ra.filt = s->filter;
            ra.n_rows = s->count; ra.n_eff = n_eff; ra.dim = s->dim; ra.metric = s->metric; ra.k = k; ra.group = rep_group;
            ra.slice_max = slice;
            ra.flagged2 = c.flagged2; ra.rep_thr = c.rep_thr; ra.rep_cnt = c.rep_cnt; ra.rep_sel = c.rep_sel;
            ra.pool_s = c.rep_s; ra.pool_row = c.rep_row; ra.pool_items = (int64_t)c.pool_items; ra.hist = c.rep_hist;
            ra.out_ids = out_ids; ra.out_scores = out_scores; ra.out_rows = out_rows; ra.out_s64 = c.out_s64; ra.status = c.status;
            ra.dstat = s->dstat; ra.gbar = c.gbar + 1; ra.err = reinterpret_cast<unsigned int*>(s->dstat + 3);
            void* args[] = {&ra};
            AVS_CUDA(cudaLaunchCooperativeKernel((const void*)repair_kernel, dim3((unsigned)(rep_ctas_per_sm[s->device & 63] * s->num_sms)),
                                                 dim3(REPAIR_THREADS), args, rep_smem, st));
            s->st_launches += 2;
        }
        s->st_last_path = 3;                 // exact master scan
        s->st_last_levels = 0;
        s->st_last_final_rows = s->count;
        AVS_CUDA(cudaEventRecord(pev, st));
        if (s->h_stats) AVS_CUDA(cudaMemcpyAsync(s->h_stats, s->dstat, 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
        return AVS_OK;
    }

    // Level schedule -> launches.  Tensor-core path: ONE persistent launch scans every level and runs the selects
    // between them (scan_gemm.cu).  Hybrid small-batch path: warp-dot dense level + its select, then one persistent
    // launch for the remaining levels.  Warp-dot path: a scan and a select launch per level.
    const int first_gemm_level = use_gemm ? 0 : (hybrid_on ? 1 : L);
    for (int l = 0; l < first_gemm_level; ++l) {
        const bool final_level = (l == L - 1);
        size_t slot = 0;
        const bool timed = final_level && timing_begin(s, st, &slot);
        for (int q0 = 0; q0 < nq; q0 += 8) AVS_CHECK(avs_launch_scan_gemv(s, q0, nq - q0 < 8 ? nq - q0 : 8, lv[l], cap, st));
        if (timed) timing_end(s, st, slot);
        SelectArgs sa = {c.cand, c.cnt, cap, c.tau, j_ranks[l], final_level ? 1 : 0, kprime, n_eff, c.topkeys, c.topn,
                         bound, c.status, lv[l].dense ? (int)(lv[l].n_visit * AVS_GROUP_ROWS) : 0, nq,
                         lv[l].dense ? c.dense_buf : nullptr, AVS_DENSE_CAP,
                         eps_used, 0};
        select_level_kernel<<<nq, nq <= 64 ? 1024 : 256, (size_t)cap * 8, st>>>(sa);
        s->st_launches++;
        AVS_CUDA(cudaGetLastError());
    }
    if (first_gemm_level < L) {
        AvsScanPlan plan;
        memset(&plan, 0, sizeof(plan));
        plan.n_levels = L - first_gemm_level;
        plan.last_is_final = 1;
        plan.nq = nq; plan.kprime = kprime; plan.cap = cap; plan.n_eff = n_eff;
        int64_t rows_scanned = 0;
        for (int l = first_gemm_level; l < L; ++l) {
            const int i = l - first_gemm_level;
            plan.lv[i] = lv[l];
            plan.j_rank[i] = j_ranks[l];
            plan.k_eps[i] = (l == L - 2 && fine_levels && s->eps_rule) ? k : 0;
            rows_scanned += lv[l].n_visit * AVS_GROUP_ROWS;
        }
        plan.tau = c.tau; plan.cand = c.cand; plan.cnt = c.cnt; plan.topkeys = c.topkeys; plan.topn = c.topn;
        plan.bound = bound; plan.status = c.status; plan.eps = eps_used;
        plan.gbar = c.gbar + 0;
        plan.err = reinterpret_cast<unsigned int*>(s->dstat + 3);
        plan.trace = s->opt_trace ? c.trace : nullptr;
        s->st_last_final_rows = rows_scanned < s->count ? rows_scanned : s->count;   // rows the timed launch scans
        size_t slot = 0;
        const bool timed = timing_begin(s, st, &slot);
        AVS_CHECK(avs_launch_scan_gemm(s, nq, plan, st));
        if (timed) timing_end(s, st, slot);
    }

    const int fin_threads = s->opt_finalize_threads > 0 ? s->opt_finalize_threads : (nq <= 64 ? 1024 : 256);
    // finalize -> wide -> repair: programmatic dependent launches (every kernel waits for its predecessor on the
    // device before it reads anything; what is saved is the launch latency between them)
    const bool pdl = s->opt_pdl != 0;
    if (fin_threads == 256)
        AVS_CUDA(avs_launch(finalize_kernel<256>, dim3(nq), dim3(256), 0, st, pdl,
                            s->master, s->ids, q, c.qnorm, s->dim, s->metric, c.topkeys, c.topn, bound, eps_used, kprime, k,
                            n_eff, s->opt_force_repair, out_ids, out_scores, out_rows, c.out_s64, c.status,
                            c.flagged, c.rep_thr, c.rep_cnt, s->dstat));
    else
        AVS_CUDA(avs_launch(finalize_kernel<1024>, dim3(nq), dim3(fin_threads), 0, st, pdl,
                            s->master, s->ids, q, c.qnorm, s->dim, s->metric, c.topkeys, c.topn, bound, eps_used, kprime, k,
                            n_eff, s->opt_force_repair, out_ids, out_scores, out_rows, c.out_s64, c.status,
                            c.flagged, c.rep_thr, c.rep_cnt, s->dstat));
    s->st_launches++;
    AVS_CUDA(avs_launch(wide_rescore_kernel, dim3(nq), dim3(1024), AVS_WIDE_MAX * sizeof(Hit), st, pdl,
                        s->master, s->ids, q, c.qnorm, s->dim, s->metric, c.cand, c.cnt, cap, c.tau, eps_used, k,
                        n_eff, s->opt_force_repair, out_ids, out_scores, out_rows, c.out_s64, c.status, c.flagged, c.flagged2, c.rep_thr,
                        c.rep_cnt, s->dstat));
    s->st_launches++;
    {
        RepairArgs ra;
        ra.master = s->master; ra.q = q; ra.qnorm = c.qnorm; ra.ids = s->ids; ra.filt = s->filter;
        ra.n_rows = s->count; ra.n_eff = n_eff; ra.dim = s->dim; ra.metric = s->metric; ra.k = k; ra.group = rep_group;
        ra.slice_max = AVS_REPAIR_CAP;
        ra.flagged2 = c.flagged2; ra.rep_thr = c.rep_thr; ra.rep_cnt = c.rep_cnt; ra.rep_sel = c.rep_sel;
        ra.pool_s = c.rep_s; ra.pool_row = c.rep_row; ra.pool_items = (int64_t)c.pool_items; ra.hist = c.rep_hist;
        ra.out_ids = out_ids; ra.out_scores = out_scores; ra.out_rows = out_rows; ra.out_s64 = c.out_s64; ra.status = c.status;
        ra.dstat = s->dstat; ra.gbar = c.gbar + 1; ra.err = reinterpret_cast<unsigned int*>(s->dstat + 3);
        // grid = what the device keeps resident (occupancy query above), like the persistent scan: its grid barriers
        // cannot wait for a CTA that never starts.  With PDL the launch cannot carry the cooperative attribute.
        const dim3 rep_grid((unsigned)(rep_ctas_per_sm[s->device & 63] * s->num_sms));
        if (pdl) AVS_CUDA(avs_launch(repair_kernel, rep_grid, dim3(REPAIR_THREADS), rep_smem, st, true, ra));
        else {
            void* args[] = {&ra};
            AVS_CUDA(cudaLaunchCooperativeKernel((const void*)repair_kernel, rep_grid, dim3(REPAIR_THREADS), args, rep_smem, st));
        }
        s->st_launches++;
    }
    AVS_CUDA(cudaEventRecord(pev, st));
    if (s->h_stats) AVS_CUDA(cudaMemcpyAsync(s->h_stats, s->dstat, 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    return AVS_OK;
}

extern "C" int avs_search(avs_store* s, const float* q, int nq, int k, int64_t* out_ids, float* out_scores,
                          int64_t* out_rows, void* stream) {
    return avs_search_local(s, q, nq, k, out_ids, out_scores, out_rows, (cudaStream_t)stream);
}

// one device block [ids | rows | scores] and one pinned host mirror: a single D2H copy brings every result back
int avs_host_staging_reserve(avs_store* s, int nq, int k) {
    AvsScratch& c = s->sc;
    if (nq <= c.host_nq_cap && k <= c.host_k_cap) return AVS_OK;
    AVS_CUDA(cudaDeviceSynchronize());
    const int nq2 = nq > c.host_nq_cap ? nq : c.host_nq_cap, k2 = k > c.host_k_cap ? k : c.host_k_cap;
    const size_t cap_items = (size_t)nq2 * k2;
    AVS_CHECK(dev_alloc(&c.h2d_q, (size_t)nq2 * s->dim));
    AVS_CHECK(dev_alloc(&c.d_ids, cap_items * 3));                 // ids, rows (int64) and scores (fp32, 8-byte slots)
    if (c.h_out) cudaFreeHost(c.h_out);
    if (c.h_q) cudaFreeHost(c.h_q);
    c.h_out = nullptr; c.h_q = nullptr;
    c.host_nq_cap = c.host_k_cap = 0;
    if (cudaHostAlloc((void**)&c.h_out, cap_items * 3 * sizeof(int64_t), cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void**)&c.h_q, (size_t)nq2 * s->dim * sizeof(float), cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        avs_set_error("out of pinned host memory for the search staging buffers");
        return AVS_E_NOMEM;
    }
    c.host_nq_cap = nq2; c.host_k_cap = k2;
    return AVS_OK;
}

extern "C" int avs_search_host(avs_store* s, const float* q_host, int nq, int k, int64_t* out_ids_host,
                               float* out_scores_host, int64_t* out_rows_host) {
    if (!s) { avs_set_error("avs_search_host: NULL store"); return AVS_E_INVALID; }
    if (nq < 0 || (nq > 0 && (!q_host || !out_ids_host || !out_scores_host))) { avs_set_error("avs_search_host: NULL buffer"); return AVS_E_INVALID; }
    if (k < 1 || k > AVS_MAX_LIMIT) { avs_set_error("avs_search: limit %d outside [1, %d]", k, AVS_MAX_LIMIT); return AVS_E_INVALID; }
    if (nq == 0) return AVS_OK;
    AVS_CUDA(cudaSetDevice(s->device));
    AVS_CHECK(avs_host_staging_reserve(s, nq, k));
    AvsScratch& c = s->sc;
    const size_t items = (size_t)nq * k;
    int64_t* d_ids = c.d_ids;
    int64_t* d_rows = c.d_ids + items;
    float* d_scores = reinterpret_cast<float*>(c.d_ids + 2 * items);
    cudaStream_t st = 0;
    // queries: from the caller's buffer directly when it is already pinned (no extra host copy), else via the pinned stage
    cudaPointerAttributes pa;
    const bool pinned = cudaPointerGetAttributes(&pa, q_host) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const float* src = q_host;
    if (!pinned) { memcpy(c.h_q, q_host, (size_t)nq * s->dim * sizeof(float)); src = c.h_q; }
    AVS_CUDA(cudaMemcpyAsync(c.h2d_q, src, (size_t)nq * s->dim * sizeof(float), cudaMemcpyHostToDevice, st));
    AVS_CHECK(avs_search_local(s, c.h2d_q, nq, k, d_ids, d_scores, d_rows, st));
    const size_t out_bytes = items * (2 * sizeof(int64_t) + sizeof(float));
    AVS_CUDA(cudaMemcpyAsync(c.h_out, c.d_ids, out_bytes, cudaMemcpyDeviceToHost, st));
    AVS_CUDA(cudaStreamSynchronize(st));
    if (s->h_stats) {   // the mirror copy of the counters was enqueued before the result copy: current after the sync
        s->st_last_uncertified = (int64_t)(s->h_stats[1] - s->seen_uncertified);
        s->seen_uncertified = s->h_stats[1];
    }
    memcpy(out_ids_host, c.h_out, items * sizeof(int64_t));
    if (out_rows_host) memcpy(out_rows_host, c.h_out + items, items * sizeof(int64_t));
    memcpy(out_scores_host, c.h_out + 2 * items, items * sizeof(float));
    return AVS_OK;
}

extern "C" int avs_set_option(avs_store* s, const char* key, int64_t value) {
    if (!s || !key) { avs_set_error("avs_set_option: NULL argument"); return AVS_E_INVALID; }
    std::string k(key);
    if (k == "scan_path") s->opt_scan_path = (int)value;
    else if (k == "oversample") s->opt_oversample = (int)value;
    else if (k == "gemm_min_batch") s->opt_gemm_min_batch = (int)value;
    else if (k == "levels_ratio") s->opt_ratio = (int)value;
    else if (k == "force_repair") s->opt_force_repair = (int)value;
    else if (k == "cta_group") s->opt_cta_group = value == 2 ? 2 : 1;
    else if (k == "p2p_merge") s->opt_p2p = value != 0;
    else if (k == "p2p_timeout_ms") s->opt_p2p_timeout_ms = value < 0 ? 0 : (long long)value;
    else if (k == "final_sigma") s->opt_final_sigma = value < 1 ? 1 : (int)value;
    else if (k == "fine_ratio") s->opt_fine_ratio = value < 2 ? 2 : (int)value;
    else if (k == "hybrid") s->opt_hybrid = value != 0;
    else if (k == "boot") s->opt_boot = value != 0;
    else if (k == "trace") s->opt_trace = value != 0;
    else if (k == "pdl") s->opt_pdl = value != 0;
    else if (k == "eps_rule") s->eps_rule = value != 0;   // normally switched on by the first exact repair (tests force it)
    else if (k == "boot2_ratio") s->opt_boot2_ratio = value < 0 ? 0 : (value > 64 ? 64 : (int)value);
    else if (k == "finalize_threads") s->opt_finalize_threads = (value == 256 || value == 512 || value == 1024) ? (int)value : 0;
    else if (k == "gemm_dense_rows") s->opt_gemm_dense_rows = value < 256 ? 256 : (value > 2048 ? 2048 : (int)(value / 256 * 256));
    else if (k == "dense_rows") s->opt_dense_rows = value < 2048 ? 2048 : (value > AVS_DENSE_CAP ? AVS_DENSE_CAP : (int)value);
    else if (k == "fine_min_batch") s->opt_fine_min_batch = value < 1 ? 1 : (int)value;
    else if (k == "coarse_sigma") s->opt_coarse_sigma = value < 1 ? 1 : (int)value;
    else if (k == "cta_group_small") s->opt_cta_group_small = value == 1 ? 1 : 2;
    else { avs_set_error("avs_set_option: unknown option '%s'", key); return AVS_E_INVALID; }
    return AVS_OK;
}

extern "C" int avs_get_stat(avs_store* s, const char* key, int64_t* out) {
    if (!s || !key || !out) { avs_set_error("avs_get_stat: NULL argument"); return AVS_E_INVALID; }
    std::string k(key);
    if (k == "kernel_launches") *out = s->st_launches;
    else if (k == "searches") *out = s->st_searches;
    else if (k == "queries") *out = s->st_queries;
    else if (k == "last_kprime") *out = s->st_last_kprime;
    else if (k == "last_levels") *out = s->st_last_levels;
    else if (k == "last_boot") *out = s->st_last_boot;
    else if (k == "last_final_rows") *out = s->st_last_final_rows;
    else if (k == "p2p_timeouts") { AVS_CUDA(cudaSetDevice(s->device)); return avs_p2p_timeouts(s, out); }
    else if (k == "exchange_us") { AVS_CUDA(cudaSetDevice(s->device)); return avs_p2p_exchange_us(s, out); }
    else if (k == "last_scan_path") *out = s->st_last_path;
    else if (k.rfind("trace:", 0) == 0) {      // phase timestamp i (ns, globaltimer) of the last persistent scan: synchronises
        const int i = atoi(k.c_str() + 6);
        if (i < 0 || i >= AVS_TRACE_SLOTS || !s->sc.trace) { avs_set_error("avs_get_stat: trace slot out of range"); return AVS_E_INVALID; }
        AVS_CUDA(cudaSetDevice(s->device));
        u64 v = 0;
        AVS_CUDA(cudaMemcpy(&v, s->sc.trace + i, sizeof(v), cudaMemcpyDeviceToHost));
        *out = (int64_t)v;
    }
    else if (k == "last_uncertified") *out = s->st_last_uncertified;   // of the last avs_search_host call; no device sync
    else if (k == "repaired_queries" || k == "uncertified_queries" || k == "wide_rescored_queries" || k == "barrier_timeouts") {
        AVS_CUDA(cudaSetDevice(s->device));
        u64 h[4];
        AVS_CUDA(cudaMemcpy(h, s->dstat, sizeof(h), cudaMemcpyDeviceToHost));
        *out = (int64_t)(k == "repaired_queries" ? h[0] : k == "uncertified_queries" ? h[1] : k == "wide_rescored_queries" ? h[2] : (h[3] & 0xFFFFFFFFull));
    } else { avs_set_error("avs_get_stat: unknown stat '%s'", key); return AVS_E_INVALID; }
    return AVS_OK;
}
