// Internal declarations shared by the translation units of libavs.so.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/avs.h"

typedef unsigned long long u64;

#define AVS_GROUP_ROWS 256      // sampling / tiling granularity of the scan (rows)
#define AVS_MAX_KPRIME 256      // largest oversampled candidate list
#define AVS_REPAIR_CAP 4096     // largest exact-repair slice per flagged query (rows tied at the k-th score beyond it: uncertified)
#define AVS_MAX_LIMIT 16384     // largest `limit` (MilvusClient's own); limits above AVS_MAX_KPRIME are served by the exact master scan
#define AVS_MAX_LEVELS 12
#define AVS_DENSE_CAP 65536      // rows of the threshold-free level of the gemv path (dense key buffer per query)
#define AVS_DENSE_MAX_NQ 64      // the dense buffer is kept for this many queries
#define AVS_TRACE_SEL (64 + AVS_MAX_LEVELS * 256 * 9)   // trace buffer: 16 phase stamps of CTA 0's level select per level start here
#define AVS_TRACE_SLOTS (AVS_TRACE_SEL + AVS_MAX_LEVELS * 16)
#define AVS_BOOT_J 8             // tensor-core path, boot level: keys kept per (row group half, query); the level's rank must not exceed it

// ---- error plumbing -------------------------------------------------------------------------
void avs_set_error(const char* fmt, ...);
#define AVS_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            avs_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                          __LINE__);                                                           \
            return AVS_E_CUDA;                                                                 \
        }                                                                                      \
    } while (0)
#define AVS_CHECK(expr)                                                                        \
    do {                                                                                       \
        int _r = (expr);                                                                       \
        if (_r != AVS_OK) return _r;                                                           \
    } while (0)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// The kernels of one search form a chain prep -> scan -> finalize -> wide -> repair.  Each is launched with the
// programmatic-stream-serialization attribute, raises `launch_dependents` first thing (its successor's CTAs may then be
// staged on whatever SM resources are free and run their own set-up) and executes `griddepcontrol.wait` - every
// thread, before any early exit and before it touches global memory a predecessor writes - which returns once the
// predecessor grid has completed and its writes are visible.  A kernel that skipped the wait could complete before
// its predecessor and break the chain's transitivity for ITS successor, hence "every thread, unconditionally".
#ifdef __CUDACC__
__device__ __forceinline__ void avs_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void avs_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t avs_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---- order-preserving keys ------------------------------------------------------------------
// key = (monotone(float score) << 32) | (0xFFFFFFFF - row): a larger key is a better
// candidate (higher score, then smaller row).  Key 0 is below every real key.
__host__ __device__ __forceinline__ uint32_t avs_f2ord(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    uint32_t b;
    memcpy(&b, &f, 4);
#endif
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__host__ __device__ __forceinline__ float avs_ord2f(uint32_t o) {
    uint32_t b = o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu);
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}
__host__ __device__ __forceinline__ u64 avs_make_key(float s, uint32_t row) {
    return ((u64)avs_f2ord(s) << 32) | (u64)(0xFFFFFFFFu - row);
}
__host__ __device__ __forceinline__ uint32_t avs_key_row(u64 k) {
    return 0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull);
}
__host__ __device__ __forceinline__ float avs_key_score(u64 k) { return avs_ord2f((uint32_t)(k >> 32)); }

// ---- one sampling level of the scan ---------------------------------------------------------
// A level visits row groups g = j*stride (j = 0..n_iter-1), skipping those with
// g % skip == 0 when skip != 0 (they were visited by an earlier, sparser level).
struct AvsLevel {
    int64_t n_iter;
    int64_t stride;
    int64_t skip;
    int64_t ratio;    // skip / stride (0 or 1: nothing skipped)
    int64_t n_visit;  // groups actually visited = n_iter - ceil(n_iter / ratio)
    int dense;   // 1: threshold-free level - every visited row is stored at slot (j*256 + column), no atomics
                 // 2: threshold-free boot level of the tensor-core scan - only the AVS_BOOT_J best keys of every half
                 //    group are stored (slot (j*2 + half)*AVS_BOOT_J + i): exact for any rank <= AVS_BOOT_J
};

// m-th row group a level visits: groups j*stride, j = 0.., leaving out every ratio-th j (seen by a sparser level)
__host__ __device__ __forceinline__ int64_t avs_level_group(const AvsLevel& lv, int64_t m) {
    const int64_t j = lv.ratio > 1 ? m + m / (lv.ratio - 1) + 1 : m;
    return j * lv.stride;
}

// Everything the persistent tensor-core scan needs for one search: the levels it scans in ONE launch and what the
// fused warp-per-query selects between them need (select_warp.cuh).
struct AvsScanPlan {
    int n_levels;                       // levels scanned by this launch
    int last_is_final;                  // its last level is the final level of the search (select -> top K' + bound)
    int nq, kprime, cap;
    int64_t n_eff;                      // rows that may be returned (filter-allowed count)
    AvsLevel lv[AVS_MAX_LEVELS];
    int j_rank[AVS_MAX_LEVELS];
    int k_eps[AVS_MAX_LEVELS];          // > 0 on the level whose select sets the LAST threshold under the eps rule
    u64* tau; u64* cand; int* cnt; u64* topkeys; int* topn; float* bound; int* status; const float* eps;
    unsigned int* gbar;                 // grid barrier counter (zeroed by prep_queries_kernel)
    unsigned int* err;                  // bumped when a bounded barrier spin gives up
    u64* trace;                         // optional [4 * AVS_MAX_LEVELS + 2] globaltimer stamps of CTA 0 (profiling option "trace")
};

struct AvsScratch {
    // query preparation
    float* qf = nullptr;          // [nq_pad, dpad] fp32, normalised for COSINE, zero padded
    __nv_bfloat16* qb = nullptr;  // [nq_pad, dpad] bf16 of qf
    double* qnorm = nullptr;      // [nq] ||q|| in float64
    float* eps_gemv = nullptr;    // [nq] certificate slack for the fp32-query scan
    float* eps_gemm = nullptr;    // [nq] certificate slack for the bf16-query scan
    // candidate collection
    u64* cand = nullptr;          // [nq, cap]
    int* cnt = nullptr;           // [nq]
    u64* tau = nullptr;           // [nq_pad] running threshold key
    u64* dense_buf = nullptr;     // [AVS_DENSE_MAX_NQ + 8, AVS_DENSE_CAP] keys of the gemv path's threshold-free level
    u64* topkeys = nullptr;       // [nq, kprime] final bf16-scan candidates, sorted desc
    int* topn = nullptr;          // [nq]
    int* status = nullptr;        // [nq] bit0 overflow, bit1 underflow, bit2 cert failed, bit3 uncertified
    double* out_s64 = nullptr;    // [nq, k] exact scores of the final hits (for the shard merge)
    // repair
    int* flagged = nullptr;       // [1 + nq] stage 1 (wide rescoring): count, then query indices
    int* flagged2 = nullptr;      // [1 + nq] stage 2 (exact scan): count, then query indices
    u64* trace = nullptr;         // [64] phase timestamps of the last persistent scan (option "trace")
    unsigned int* gbar = nullptr; // [4] grid-barrier counters of the persistent kernels (scan, repair), zeroed by prep
    double* rep_s = nullptr;      // [pool_items] exact-repair pool: scores ...
    uint32_t* rep_row = nullptr;  // [pool_items] ... and rows, cut into one slice per flagged query
    int* rep_cnt = nullptr;       // [nq] rows collected per flagged query
    double* rep_thr = nullptr;    // [nq] lower bound of the k-th best exact score per flagged query (-inf: none)
    int* rep_sel = nullptr;       // [nq, 2] threshold search: histogram bin of the k-th best, rows above it
    unsigned int* rep_hist = nullptr;  // [8, 4096] histogram of the threshold search
    size_t pool_items = 0;
    // sharded search
    void* gather_send = nullptr;  // [nq, k] (f64 score, i64 id)
    void* gather_recv = nullptr;  // [world, nq, k]
    // host staging for avs_search_host
    float* h2d_q = nullptr;       // device copy of the host queries
    int64_t* d_ids = nullptr;     // device result block [ids | rows | scores]
    int64_t* h_out = nullptr;     // pinned mirror of the result block
    float* h_q = nullptr;         // pinned stage for pageable host queries
    // sizes the buffers were allocated for
    size_t gather_items = 0;
    int nq_cap = 0, kprime_cap = 0, cap_cap = 0, k_cap = 0, world_cap = 0, host_nq_cap = 0, host_k_cap = 0;
};

struct avs_store {
    int device = 0;
    int dim = 0;
    int dpad = 0;             // dim rounded up to 64 (one 128-byte swizzle atom of bf16)
    int metric = 0;
    int64_t capacity = 0;     // rows allocated (multiple of AVS_GROUP_ROWS)
    int64_t count = 0;
    float* master = nullptr;          // [capacity, dim] fp32 rows exactly as inserted
    __nv_bfloat16* xb = nullptr;      // [capacity, dpad] bf16 scan copy (unit rows for COSINE)
    float* inv_norm = nullptr;        // [capacity] 1/||x|| (0 for zero rows)
    int64_t* ids = nullptr;           // [capacity]
    uint32_t* filter = nullptr;       // optional row bitmap (bit r set = row r may be returned), set by avs_set_filter
    int64_t filter_allowed = 0;       // number of set bits
    size_t filter_words = 0;          // allocated words
    float* gstat = nullptr;           // [4] device scalars: r_max, xnorm_max (non-negative, atomicMax on bits)
    unsigned long long* dstat = nullptr;  // [8] device counters: repaired, uncertified, wide-rescored
    unsigned long long seen_uncertified = 0;   // h_stats[1] at the end of the previous avs_search_host call
    int64_t st_last_uncertified = 0;           // queries of the last avs_search_host call whose top-k could not be proven
    unsigned long long* h_stats = nullptr;  // pinned host mirror of dstat, refreshed asynchronously after every search
    unsigned long long seen_repaired = 0;
    size_t rep_smem_seen = 0;         // shared-memory size the repair kernel's occupancy was last queried for
    bool eps_rule = false;            // set once an exact repair was needed: the last threshold then honours eps
    AvsScratch sc;
    int num_sms = 148;
    // options
    int opt_scan_path = 0, opt_oversample = 0, opt_gemm_min_batch = 1, opt_ratio = 32, opt_force_repair = 0,
        opt_cta_group = 2;
    // stats
    int64_t st_launches = 0, st_searches = 0, st_queries = 0, st_last_final_rows = 0;
    int st_last_kprime = 0, st_last_levels = 0, st_last_path = 0, st_last_boot = 0;
    // scan timing hook
    bool timing = false;
    std::vector<cudaEvent_t> tev;   // event pairs around the dominant scan launches
    size_t tev_used = 0;
    // gemm path (tensor maps etc.)
    void* gemm_state = nullptr;
    // multi-GPU
    void* nccl_comm = nullptr;
    void* p2p_state = nullptr;       // peer-memory exchange regions (comm.cu)
    int opt_p2p = 1;
    long long opt_p2p_timeout_ms = 0;  // wall-clock bound of a wait for a peer in the exchange kernels (0: default, 120 s)
    int opt_final_sigma = 2;         // expected survivors of the last level = K' + sigma * sqrt(K' * ratio)
    int opt_fine_min_batch = 129;    // tensor-core path: batches from here on use the fine (x4) dense-end schedule
    int opt_coarse_sigma = 3;        // same margin for the coarse (x32) schedule of the tensor-core path (gemv: 8)
    int opt_cta_group_small = 1;     // CTA-group size for batches of at most 128 queries (1: M = 128, half the MMA work)
    int opt_dense_rows = AVS_DENSE_CAP;   // rows of the gemv path's threshold-free level (<= AVS_DENSE_CAP)
    int opt_boot = 1;                // tensor-core path: threshold-free level as per-thread top-J lists (AvsLevel::dense == 2)
    int opt_hybrid = 1;              // auto mode, <= 8 queries: gemv dense level, tensor-core scan for the later levels
    int opt_gemm_dense_rows = 2048;  // rows of the tensor-core path's threshold-free level (inside the candidate buffer)
    int opt_finalize_threads = 0;    // 0: 1024 threads per query up to 64 queries, 256 beyond
    int opt_trace = 0;               // record per-level phase timestamps inside the persistent scan kernel
    int opt_fine_ratio = 4;          // stride ratio of the dense-end levels of the tensor-core path
    int opt_boot2_ratio = 0;         // compute-bound schedule: boot level directly in front of the final one up to this stride ratio (0: never)
    int opt_pdl = 1;                 // programmatic dependent launch between the kernels of a search
    int rank = 0, world = 1;
};

// ---- host entry points implemented across translation units ----------------------------------
int avs_launch_scan_gemv(avs_store* s, int q0, int nq, const AvsLevel& lv, int cap, cudaStream_t st);
int avs_launch_scan_gemm(avs_store* s, int nq, const AvsScanPlan& plan, cudaStream_t st);
void avs_gemm_state_free(avs_store* s);
int avs_search_local(avs_store* s, const float* q, int nq, int k, int64_t* out_ids, float* out_scores,
                     int64_t* out_rows, cudaStream_t st);
int avs_scratch_reserve(avs_store* s, int nq, int kprime, int cap, int k);
void avs_scratch_free(avs_store* s);
void avs_comm_free(avs_store* s);
int avs_p2p_timeouts(avs_store* s, int64_t* out);
int avs_p2p_exchange_us(avs_store* s, int64_t* out);
void avs_p2p_timing_reset(avs_store* s);
int avs_host_staging_reserve(avs_store* s, int nq, int k);
