// K3: tensor-core scan for large query batches — a dense Q x X^T contraction on tcgen05 with the
// top-k selection fused into the TMEM epilogue (no [B, N] score matrix ever reaches HBM).
//
//   A (M) = 128 queries per CTA (256 per CTA pair), bf16, K-major, TMA -> 128B-swizzled smem
//   B (N) = 256 database rows per tile (each CTA of a pair loads 128 of them)
//   D     = fp32 accumulators in TMEM: lane = query, column = database row, 2 x 256 columns
//           (double buffered: the MMA of tile t+1 overlaps the epilogue of tile t)
//   roles = warp 0 TMA producer | warp 1 tcgen05.mma issuer (one elected thread, leader CTA)
//           | warp 2 TMEM allocator | warps 4-11 epilogue (two warps per TMEM lane quarter)
//   epilogue: thread = (query lane, column half).  tcgen05.ld.32x32b.x32 (double buffered) -> 4 group
//           maxima -> ONE compare with the query's threshold (cached in smem); a qualifying chunk builds
//           the bit mask of its survivors (row filter = one bitmap word per chunk), stashes their 64-bit
//           keys in shared memory, and reserves their slots in the query's candidate buffer with one
//           atomicAdd per thread-tile whose result is only consumed a tile later.
//           The threshold-free (sparsest) level stores every score densely instead, no atomics.
// Work split: unit = (visited row group, query block of 128*CG queries); units are dealt round-robin to
// the CTAs (pairs), so the pairs that run side by side sweep the query blocks over the same database
// tile: one HBM read, the rest from L2.
// Algorithmic flops per launch: 2 * nq_pad * rows_visited * D_pad.
//
// ONE launch scans every level of a search (device-side level loop).  The TMA producer and the MMA issuer run
// straight through all levels - their tiles do not depend on the thresholds, so the first tiles of level l+1 are
// already in shared memory / TMEM while level l is being selected.  Only the epilogue warps synchronise:
//   flush survivors -> grid barrier -> every epilogue warp of the grid selects queries in turn (select_warp.cuh:
//   rank-j key = next threshold, survivors compacted) -> grid barrier -> next level with the new thresholds.
// The grid never exceeds what the device keeps resident (one CTA per SM), so the barrier cannot wait for a CTA that
// has not started; its spin is bounded all the same and a give-up is reported through `err`.
#include <cuda.h>

#include <type_traits>

#include "avs_internal.h"
#include "select_warp.cuh"

namespace {

constexpr int BLOCK_M = 128;        // queries per CTA
constexpr int BLOCK_N = 256;        // database rows per tile (== AVS_GROUP_ROWS)
constexpr int BLOCK_K = 64;         // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 384;      // 4 control warps + 8 epilogue warps
constexpr int TAU_CACHE = 2048;         // thresholds (as fp32 scores) of the queries THIS CTA sweeps (128 per query block), cached in smem
constexpr int STASH = 4;                // per-thread survivor stash (keys) between slot reservations
constexpr int RAW = 2;                  // per-thread raw stash: qualifying 8-score groups of the current tile, expanded after
                                        // the accumulator has been handed back to the MMA
// per-warp share of the stash: [32 x STASH keys | 32 x RAW x 8 scores | 32 x RAW group codes]; the dense level's transpose
// and the level select use its first 2 KB as scratch
constexpr int WARP_STASH_BYTES = 32 * STASH * 8 + 32 * RAW * 32 + 32 * RAW * 2;
static_assert(WARP_STASH_BYTES >= 2048 && WARP_STASH_BYTES % 16 == 0, "2 KB of scratch per warp, 16-byte aligned raw groups");
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr uint64_t HINT_EVICT_NORMAL = 0x1000000000000000ull;
constexpr uint64_t HINT_EVICT_LAST = 0x14F0000000000000ull;

template <int CG> struct Cfg {
    static constexpr int LOAD_N = BLOCK_N / CG;                  // X rows loaded by each CTA
    static constexpr int B_STAGE_BYTES = LOAD_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = CG == 1 ? 4 : 6;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + TAU_CACHE * 4 + 8 * WARP_STASH_BYTES;
    static_assert(SMEM_BYTES <= 232448, "227 KB of shared memory per CTA");
};

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_local, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar_local), "r"(cta));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int CG>
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, uint64_t hint) {
    if constexpr (CG == 1) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                     ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(hint) : "memory");
    } else {
        // completion is signalled on the LEADER CTA's barrier (peer bit cleared), as CUTLASS SM100_TMA_2SM_LOAD does
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                     ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(hint) : "memory");
    }
}

template <int CG>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 1) {
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    }
}

// tcgen05.commit: arrive on an mbarrier once every MMA issued so far by this thread has retired.
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if constexpr (CG == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    } else {
        const uint16_t mask = 3;  // the barrier at the same offset in both CTAs of the pair
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"(mask) : "memory");
    }
}

// K-major, 128-byte swizzle: rows are 128 B apart inside an 8-row atom, atoms 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address, 16-byte units
    d |= (uint64_t)0 << 16;                             // leading byte offset: unused for swizzled K-major
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset between 8-row atoms
    d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- boot level: the 8 largest of a thread's 128 scores with sorting networks (no data-dependent control flow) ----
#define AVS_CEX(a, b) { const float hi_ = fmaxf(a, b), lo_ = fminf(a, b); a = hi_; b = lo_; }
// optimal 19-comparator network, descending
__device__ __forceinline__ void boot_sort8(float (&x)[8]) {
    AVS_CEX(x[0], x[2]) AVS_CEX(x[1], x[3]) AVS_CEX(x[4], x[6]) AVS_CEX(x[5], x[7])
    AVS_CEX(x[0], x[4]) AVS_CEX(x[1], x[5]) AVS_CEX(x[2], x[6]) AVS_CEX(x[3], x[7])
    AVS_CEX(x[0], x[1]) AVS_CEX(x[2], x[3]) AVS_CEX(x[4], x[5]) AVS_CEX(x[6], x[7])
    AVS_CEX(x[2], x[4]) AVS_CEX(x[3], x[5])
    AVS_CEX(x[1], x[4]) AVS_CEX(x[3], x[6])
    AVS_CEX(x[1], x[2]) AVS_CEX(x[3], x[4]) AVS_CEX(x[5], x[6])
}
// a, b sorted descending -> a = the 8 largest of both, sorted descending (max against the reversed list, bitonic merge)
__device__ __forceinline__ void boot_merge8(float (&a)[8], const float (&b)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i], b[7 - i]);
    AVS_CEX(a[0], a[4]) AVS_CEX(a[1], a[5]) AVS_CEX(a[2], a[6]) AVS_CEX(a[3], a[7])
    AVS_CEX(a[0], a[2]) AVS_CEX(a[1], a[3]) AVS_CEX(a[4], a[6]) AVS_CEX(a[5], a[7])
    AVS_CEX(a[0], a[1]) AVS_CEX(a[2], a[3]) AVS_CEX(a[4], a[5]) AVS_CEX(a[6], a[7])
}
// folds 32 scores into the running sorted top-8; `allow` masks rows that must not count (padding, row filter)
__device__ __forceinline__ void boot_fold32(const uint32_t (&v)[32], uint32_t allow, bool masked, float (&top)[8]) {
    float g[4][8];
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float sc = __uint_as_float(v[8 * gi + i]) + 0.0f;   // -0 -> +0: equal values must have equal keys
            g[gi][i] = (!masked || ((allow >> (8 * gi + i)) & 1u)) ? sc : -INFINITY;
        }
        boot_sort8(g[gi]);
    }
    boot_merge8(g[0], g[1]);
    boot_merge8(g[2], g[3]);
    boot_merge8(g[0], g[2]);
    boot_merge8(top, g[0]);
}
// second pass: the (score, column) pairs of the scores above `v8` go to slots 0.., the first `ties_left` scores equal to
// it (column order) to slots 7, 6, ..; `ent` = 8 slots of this thread in shared memory
__device__ __forceinline__ void boot_emit32(const uint32_t (&v)[32], uint32_t allow, bool masked, int col0, float v8, int& pf, int& pb,
                                            int& ties_left, uint2* ent) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float sc = __uint_as_float(v[i]) + 0.0f;
        const bool ok = !masked || ((allow >> i) & 1u);
        const bool gt = ok && sc > v8;
        const bool tie = ok && sc == v8 && ties_left > 0;
        if (gt || tie) ent[gt ? pf : pb] = make_uint2(__float_as_uint(sc), (uint32_t)(col0 + i));
        pf += gt ? 1 : 0;
        pb -= tie ? 1 : 0;
        ties_left -= tie ? 1 : 0;
    }
}

struct PipeState {
    uint32_t stage = 0, phase = 0;
    template <int N> __device__ __forceinline__ void advance() {
        if (++stage == N) { stage = 0; phase ^= 1; }
    }
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps

// Grid-wide barrier among the epilogue groups of all CTAs; called by every epilogue thread.
__device__ __forceinline__ void grid_barrier_epi(unsigned int* gbar, unsigned int& epoch, unsigned int* err) {
    __threadfence();
    epi_sync();
    epoch += 1;
    if (threadIdx.x == 128) {
        const unsigned int target = epoch * gridDim.x;
        atomicAdd(gbar, 1u);
        long long spins = 0;
        while (ld_acquire_u32(gbar) < target) {
            __nanosleep(20);
            if (++spins > (1ll << 22)) { atomicAdd(err, 1u); break; }   // ~2 s: a CTA is missing; results are flagged, the GPU is not hung
        }
        __threadfence();
    }
    epi_sync();
}

// Expands the parked groups of one thread (see the epilogue) into keys in its key stash; returns the new stash fill.
// Survivors beyond the stash (duplicate-heavy data, dense accepts of the sparse levels) reserve their slots one by one
// and go straight to the candidate buffer.  Out of line on purpose: five call sites, all off the hot path.
__device__ __noinline__ int drain_raw(const float4* raw, const uint16_t* code_of, int n_raw, float tau_f,
                                      const uint32_t* __restrict__ filt, int64_t row0, int valid_cols, u64* stash, int n_stash,
                                      int* cnt_q, u64* cand_q, int cap) {
    for (int e = 0; e < n_raw; ++e) {
        const float4 lo4 = raw[2 * e], hi4 = raw[2 * e + 1];
        const int code = code_of[e];
        const int colb = (code >> 2) * 32 + (code & 3) * 8;          // first column of the group inside the tile half
        const float g[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
        uint32_t allow = 0xFFu;
        if (filt) allow = (filt[(row0 + colb) >> 5] >> (colb & 31)) & 0xFFu;   // row filter: 8 bits of the chunk's bitmap word
        const int vc = valid_cols - colb;                            // padding rows of the store's last group
        allow = vc >= 8 ? allow : (vc <= 0 ? 0u : (allow & ((1u << vc) - 1)));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (g[i] >= tau_f && ((allow >> i) & 1u)) {
                const u64 key = avs_make_key(g[i], (uint32_t)(row0 + colb + i));
                if (n_stash < STASH) stash[n_stash++] = key;
                else { const int pos = atomicAdd(cnt_q, 1); if (pos < cap) cand_q[pos] = key; }
            }
        }
    }
    return n_stash;
}

template <int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
scan_gemm_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x,
                 int64_t n_rows, int n_qblocks, int num_k_blocks, const __grid_constant__ AvsScanPlan plan,
                 const uint32_t* __restrict__ filt) {
    using C = Cfg<CG>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* full_bar = bars;                       // [STAGES] TMA bytes landed (leader CTA's copy is the live one)
    uint64_t* empty_bar = bars + C::STAGES;          // [STAGES] MMA has consumed the stage
    uint64_t* tfull_bar = bars + 2 * C::STAGES;      // [2] accumulator ready for the epilogue
    uint64_t* tempty_bar = bars + 2 * C::STAGES + 2; // [2] accumulator drained (leader CTA's copy is live)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
    float* tau_smem = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + 256);
    u64* stash_smem = reinterpret_cast<u64*>(tau_smem + TAU_CACHE);
    const int n_tau = n_qblocks * BLOCK_M;           // queries this CTA sweeps: 128 of every query block
    const bool tau_cached = n_tau <= TAU_CACHE;
    const u64* __restrict__ tau = plan.tau;
    u64* __restrict__ cand = plan.cand;
    int* __restrict__ cnt = plan.cnt;
    const int cap = plan.cap;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = CG == 1 ? 0u : cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int cluster_id = blockIdx.x / CG, n_clusters = gridDim.x / CG;
    // work unit = (visited row group, query block); units are dealt round-robin to the CTAs (pairs), so the
    // CTAs that run side by side sweep the query blocks over the same database tile: one HBM read, L2 re-use

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < C::STAGES; ++i) {
            mbar_init(smem_u32(full_bar + i), 1);    // leader's arrive.expect_tx covers the bytes of both CTAs
            mbar_init(smem_u32(empty_bar + i), 1);   // one tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(tfull_bar + i), 1);       // one tcgen05.commit
            mbar_init(smem_u32(tempty_bar + i), 8 * CG); // one arrive per epilogue warp of every CTA
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if constexpr (CG == 1) __syncthreads(); else cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // programmatic dependent launch: everything above ran while prep_queries was still preparing the batch; from here on
    // the kernel reads what prep wrote (queries, thresholds, counters).  finalize may be staged right away.
    avs_pdl_trigger();
    avs_pdl_wait();

    if (warp == 0) {
        // ===== TMA producer (one lane) =====
        // Tile coordinates are computed once per TILE (two 64-bit divisions), never per k-block: the lane has ~700 cycles
        // per k-block on the compute-bound path, and a flat (level, tile, k-block) cursor that redid the divisions every
        // k-block made this thread the bottleneck of the whole kernel (batch 1024: 1.13 -> 1.47 ms, profiles/r02/c26).
        if (lane == 0) {
            PipeState ps;
            for (int l = 0; l < plan.n_levels; ++l) {
                const AvsLevel& lv = plan.lv[l];
                const int64_t n_tiles = lv.n_visit * n_qblocks;
                for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
                    const int64_t m = t / n_qblocks;
                    const int qb = (int)(t - m * n_qblocks);
                    const int64_t g = avs_level_group(lv, m);
                    const int x_row = (int)(g * BLOCK_N + cta_rank * C::LOAD_N);
                    const int q_row = (qb * CG + (int)cta_rank) * BLOCK_M;
                    for (int kb = 0; kb < num_k_blocks; ++kb) {
                        mbar_wait(smem_u32(empty_bar + ps.stage), ps.phase ^ 1);
                        const uint32_t fb = smem_u32(full_bar + ps.stage);
                        uint8_t* sa = smem + ps.stage * C::STAGE_BYTES;
                        if (CG == 1 || leader) mbar_arrive_expect_tx(fb, C::STAGE_BYTES * CG);
                        tma_load_2d<CG>(&map_q, fb, smem_u32(sa), kb * BLOCK_K, q_row, HINT_EVICT_LAST);
                        tma_load_2d<CG>(&map_x, fb, smem_u32(sa + A_STAGE_BYTES), kb * BLOCK_K, x_row, HINT_EVICT_NORMAL);
                        ps.template advance<C::STAGES>();
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA, one lane) =====
        if (leader && lane == 0) {
            // instruction descriptor: D fp32, A/B bf16, both K-major, N = 256, M = 128 * CG
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                                   ((uint32_t)((BLOCK_M * CG) >> 4) << 24);
            PipeState ps;
            uint32_t acc = 0, acc_phase = 0;
            for (int l = 0; l < plan.n_levels; ++l) {
                const int64_t n_tiles = plan.lv[l].n_visit * n_qblocks;
                for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
                    mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                    for (int kb = 0; kb < num_k_blocks; ++kb) {
                        mbar_wait(smem_u32(full_bar + ps.stage), ps.phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + ps.stage * C::STAGE_BYTES);
                        const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + A_STAGE_BYTES);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            // advance 32 bytes (16 bf16) inside the 128-byte swizzle row: +2 in 16-byte units
                            umma_bf16<CG>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit<CG>(smem_u32(empty_bar + ps.stage));   // frees the smem stage (both CTAs)
                        ps.template advance<C::STAGES>();
                    }
                    umma_commit<CG>(smem_u32(tfull_bar + acc));            // accumulator complete (both CTAs)
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: 8 warps; thread = (query lane, column half) =====
        const int ew = warp & 3;                       // TMEM lane quarter this warp may touch
        const int half = (warp - 4) >> 2;              // columns [half*128, half*128 + 128) of the tile
        uint32_t acc = 0, acc_phase = 0;
        uint8_t* const warp_stash = reinterpret_cast<uint8_t*>(stash_smem) + (size_t)(warp - 4) * WARP_STASH_BYTES;
        u64* const my_stash = reinterpret_cast<u64*>(warp_stash) + lane * STASH;
        float4* const my_raw = reinterpret_cast<float4*>(warp_stash + 32 * STASH * 8) + lane * RAW * 2;
        uint16_t* const my_code = reinterpret_cast<uint16_t*>(warp_stash + 32 * STASH * 8 + 32 * RAW * 32) + lane * RAW;
        // survivors of a tile wait in this thread's shared-memory stash; their slots in the query's candidate
        // buffer are reserved by ONE atomicAdd issued at the end of the tile and consumed a tile later, so the
        // L2 round trip of the atomic never stalls the warp
        u64* pend_dst = nullptr;
        int pend_n = 0, pend_pos = 0;
        auto flush_pending = [&]() {
            for (int i = 0; i < pend_n; ++i)
                if (pend_pos + i < cap) pend_dst[pend_pos + i] = my_stash[i];
            pend_n = 0;
        };
        unsigned int epoch = 0;
        const bool tracer = plan.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 128;
        auto stamp = [&](int slot) {
            if (tracer) { u64 t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); plan.trace[slot] = t; }
        };
        stamp(0);
        for (int l = 0; l < plan.n_levels; ++l) {
        const AvsLevel& lv = plan.lv[l];
        const int64_t n_tiles = lv.n_visit * n_qblocks;
        if (tau_cached) {                              // this level's thresholds (the previous select wrote them)
            for (int i = threadIdx.x - 128; i < n_tau; i += 256) {
                const u64 tk = tau[((i / BLOCK_M) * CG + (int)cta_rank) * BLOCK_M + (i % BLOCK_M)];
                tau_smem[i] = tk == 0ull ? -INFINITY : avs_key_score(tk);   // NaN for padding slots (key ~0): never accepts
            }
            epi_sync();
        }
        // the tile loop exists twice - threshold-free (dense) level and thresholded levels - so that the dense level's
        // store code does not sit inside the hot loop of the thresholded levels
        auto run_tiles = [&](auto mode_tag) {
        constexpr int MODE = decltype(mode_tag)::value;      // 0 thresholded, 1 dense (every key stored), 2 boot (top-J per thread)
        constexpr bool DENSE = MODE == 1;
        constexpr bool BOOT = MODE == 2;                     // boot level, sorting networks: any number of live lanes
        constexpr bool BOOT_FEW = MODE == 3;                 // boot level, sorted insertion: a handful of live lanes (<= 8 queries)
        for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
            const int64_t m = t / n_qblocks;
            const int qb = (int)(t - m * n_qblocks);
            const int64_t row0 = avs_level_group(lv, m) * BLOCK_N + half * 128;
            const bool has_pad = row0 + 128 > n_rows;  // only the last group of the store can hold padding rows
            const int valid_cols = has_pad ? (int)(n_rows > row0 ? n_rows - row0 : 0) : 128;
            const int q = (qb * CG + (int)cta_rank) * BLOCK_M + ew * 32 + lane;
            float tau_f;
            if (tau_cached) tau_f = tau_smem[qb * BLOCK_M + ew * 32 + lane];
            else { const u64 tau_k = tau[q]; tau_f = tau_k == 0ull ? -INFINITY : avs_key_score(tau_k); }   // NaN for padding slots
            u64* const my_cand = cand + (size_t)q * cap;
            flush_pending();
            int n_stash = 0, n_raw = 0;
            auto reserve = [&]() {
                pend_n = n_stash;
                pend_dst = my_cand;
                pend_pos = atomicAdd(cnt + q, n_stash);   // result is first touched by flush_pending()
                n_stash = 0;
            };
            // Accepts are rare and hit ONE lane of a warp: whatever that lane does, the other 31 wait, and the tile's
            // accumulator is not handed back to the MMA before the slowest of the 16 epilogue warps is through
            // (measured: ~3 k cycles per accept when the survivors were expanded in place).  So the hot path only parks
            // a qualifying group of 8 scores in the thread's raw stash (three shared-memory stores); drain_raw() turns
            // the parked groups into keys AFTER the hand-back, off the MMA's critical path.
            auto drain = [&]() {
                n_stash = drain_raw(my_raw, my_code, n_raw, tau_f, filt, row0, valid_cols, my_stash, n_stash, cnt + q, my_cand, cap);
                n_raw = 0;
            };
            // 32 scores of this thread's query.  Dense level (threshold-free, the sparsest level): every score goes
            // to its own slot, no atomics.  Otherwise 4 independent group maxima (short dependency chains) and ONE
            // compare against the threshold; only a qualifying chunk builds the bit mask of its survivors.
            auto process = [&](const uint32_t (&v)[32], int col0) {
                if constexpr (DENSE) {
                    // A lane owns one query's 32 scores, and the query's keys are contiguous in memory: stored lane by
                    // lane, every store instruction would touch 32 sectors for 8 bytes each (1 K LSU transactions per
                    // warp and chunk; the 2 048-row level cost 36 us that way).  So the warp transposes 32 queries x 8
                    // keys at a time through its 2 KB of the (idle) survivor stash, XOR-swizzled so that neither side
                    // has bank conflicts, and stores 4 queries x 64 contiguous bytes per instruction.
                    const uint32_t allow = filt ? filt[(row0 + col0) >> 5] : ~0u;   // the chunk's 32 rows = one bitmap word
                    const int live_lanes = plan.nq - (q - lane);                    // real queries among this warp's 32 lanes
                    if (live_lanes <= 8) {
                        // a handful of queries (the reference's batch-1 search): the live lanes store their 32 keys
                        // themselves - 32 narrow stores instead of the 16-step transpose through shared memory
                        if (lane < live_lanes) {
                            u64* const gq = cand + (size_t)q * cap + (size_t)m * BLOCK_N + half * 128 + col0;
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                gq[i] = (col0 + i < valid_cols && ((allow >> i) & 1u))
                                            ? avs_make_key(__uint_as_float(v[i]), (uint32_t)(row0 + col0 + i)) : 0ull;
                        }
                        return;
                    }
                    u64* const stage = reinterpret_cast<u64*>(warp_stash);
                    const int rq = lane >> 3, rc = lane & 7;                       // reader role: query 4j + rq, key rc of the slice
                    u64* const gbase = cand + (size_t)(q - lane) * cap + (size_t)m * BLOCK_N + half * 128 + col0;
                    const bool live = q < plan.nq;                                 // padding query slots keep empty keys
#pragma unroll
                    for (int sl = 0; sl < 4; ++sl) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const int i = 8 * sl + c;
                            const u64 key = (col0 + i < valid_cols && live && ((allow >> i) & 1u))
                                                ? avs_make_key(__uint_as_float(v[i]), (uint32_t)(row0 + col0 + i)) : 0ull;
                            stage[lane * 8 + (c ^ ((lane >> 1) & 7))] = key;
                        }
                        __syncwarp();
                        u64* gp = gbase + (size_t)rq * cap + 8 * sl + rc;
#pragma unroll 4
                        for (int r = rq; r < 32; r += 4, gp += (size_t)4 * cap)   // kept rolled: the kernel sits at its register limit
                            *gp = stage[r * 8 + (rc ^ ((r >> 1) & 7))];
                        __syncwarp();
                    }
                    return;
                } else {
                float gm[4];
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                    const float a = fmaxf(fmaxf(__uint_as_float(v[8 * gi]), __uint_as_float(v[8 * gi + 1])), __uint_as_float(v[8 * gi + 2]));
                    const float b = fmaxf(fmaxf(__uint_as_float(v[8 * gi + 3]), __uint_as_float(v[8 * gi + 4])), __uint_as_float(v[8 * gi + 5]));
                    gm[gi] = fmaxf(fmaxf(__uint_as_float(v[8 * gi + 6]), __uint_as_float(v[8 * gi + 7])), fmaxf(a, b));
                }
                const float mx = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
                if (mx >= tau_f) {
#pragma unroll
                    for (int gi = 0; gi < 4; ++gi) {
                        if (gm[gi] >= tau_f) {             // park the group: 8 scores + where they sit
                            if (n_raw == RAW) drain();     // a third qualifying group in one tile (dense accepts of the sparse levels)
                            my_raw[2 * n_raw] = make_float4(__uint_as_float(v[8 * gi]), __uint_as_float(v[8 * gi + 1]),
                                                            __uint_as_float(v[8 * gi + 2]), __uint_as_float(v[8 * gi + 3]));
                            my_raw[2 * n_raw + 1] = make_float4(__uint_as_float(v[8 * gi + 4]), __uint_as_float(v[8 * gi + 5]),
                                                                __uint_as_float(v[8 * gi + 6]), __uint_as_float(v[8 * gi + 7]));
                            my_code[n_raw] = (uint16_t)((col0 >> 5) * 4 + gi);
                            ++n_raw;
                        }
                    }
                }
                }
            };
            mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
            tc_fence_after();
            const uint32_t t_base = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BLOCK_N + half * 128;
            if constexpr (BOOT_FEW) {
                // Boot level for at most 8 queries: one or two warps of the CTA have a live lane at all, and a lone lane
                // inserts into its sorted top-J list only when a score beats the list's last entry (~30 times per tile),
                // so the straightforward rolled loop costs a few microseconds where the network version costs ten.
                // Strict compare: equal scores keep column order (smaller row first = larger key).
                const bool live = q < plan.nq;
                float bs[AVS_BOOT_J];
                int bc[AVS_BOOT_J];
#pragma unroll
                for (int i = 0; i < AVS_BOOT_J; ++i) { bs[i] = -INFINITY; bc[i] = -1; }
#pragma unroll 1
                for (int c8 = 0; c8 < 128; c8 += 8) {
                    uint32_t v8[8];
                    __syncwarp();
                    tmem_ld8(t_base + c8, v8);
                    tmem_ld_wait();
                    if (live) {
                        uint32_t allow = 0xFFu;
                        if (filt) allow = (filt[(row0 + c8) >> 5] >> (c8 & 31)) & 0xFFu;
                        const int vc = valid_cols - c8;
                        allow = vc >= 8 ? allow : (vc <= 0 ? 0u : (allow & ((1u << vc) - 1)));
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float sc = __uint_as_float(v8[i]) + 0.0f;      // -0 -> +0: equal values must have equal keys
                            if (sc > bs[AVS_BOOT_J - 1] && ((allow >> i) & 1u)) {
#pragma unroll
                                for (int p = AVS_BOOT_J - 1; p > 0; --p) {
                                    const bool up = sc > bs[p - 1];
                                    const bool here = !up && sc > bs[p];
                                    bs[p] = up ? bs[p - 1] : (here ? sc : bs[p]);
                                    bc[p] = up ? bc[p - 1] : (here ? c8 + i : bc[p]);
                                }
                                if (sc > bs[0]) { bs[0] = sc; bc[0] = c8 + i; }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 1 || leader) mbar_arrive(smem_u32(tempty_bar + acc));
                    else mbar_arrive_remote(smem_u32(tempty_bar + acc), 0);
                }
                if (live) {   // slot block (group ordinal m, tile half): BOOT_J keys, best first, 0 = empty
                    ulonglong2* dst = reinterpret_cast<ulonglong2*>(my_cand + ((size_t)m * 2 + half) * AVS_BOOT_J);
#pragma unroll
                    for (int i = 0; i < AVS_BOOT_J; i += 2) {
                        ulonglong2 kk;
                        kk.x = bc[i] >= 0 ? avs_make_key(bs[i], (uint32_t)(row0 + bc[i])) : 0ull;
                        kk.y = bc[i + 1] >= 0 ? avs_make_key(bs[i + 1], (uint32_t)(row0 + bc[i + 1])) : 0ull;
                        dst[i >> 1] = kk;
                    }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                continue;
            }
            if constexpr (BOOT) {
                // Threshold-free level without a dense store: the rank-j key (j <= BOOT_J) of everything the level visits
                // is the rank-j key of the union of the per-thread top-J sets, and every row at or above it is in one of
                // them.  Pass 1 folds the thread's 128 scores into their 8 largest VALUES with sorting networks (fmax /
                // fmin only: no divergence however many lanes are live); pass 2 reads the accumulator again and keeps the
                // rows above the 8th value plus as many rows equal to it (column order = key order) as complete the eight.
                const bool warp_live = q - lane < plan.nq;        // warps without a real query only hand the accumulator back
                const bool masked = has_pad || filt != nullptr;
                uint2* const ent = reinterpret_cast<uint2*>(my_raw);          // 8 (score bits, column) slots of this thread
                if (warp_live) {
                    uint32_t allow[4] = {~0u, ~0u, ~0u, ~0u};
                    if (masked) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            uint32_t a = filt ? filt[(row0 + 32 * c) >> 5] : ~0u;
                            const int vc = valid_cols - 32 * c;
                            allow[c] = vc >= 32 ? a : (vc <= 0 ? 0u : (a & ((1u << vc) - 1)));
                        }
                    }
                    float top[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { top[i] = -INFINITY; ent[i] = make_uint2(0u, 0xFFFFFFFFu); }
                    uint32_t va[32], vb[32];
                    __syncwarp();
                    tmem_ld32(t_base, va);
                    tmem_ld_wait();
                    tmem_ld32(t_base + 32, vb);
                    boot_fold32(va, allow[0], masked, top);
                    __syncwarp();
                    tmem_ld_wait();
                    tmem_ld32(t_base + 64, va);
                    boot_fold32(vb, allow[1], masked, top);
                    __syncwarp();
                    tmem_ld_wait();
                    tmem_ld32(t_base + 96, vb);
                    boot_fold32(va, allow[2], masked, top);
                    __syncwarp();
                    tmem_ld_wait();
                    tmem_ld32(t_base, va);
                    boot_fold32(vb, allow[3], masked, top);
                    const float v8 = top[7];
                    int n_gt = 0;
#pragma unroll
                    for (int i = 0; i < 7; ++i) n_gt += top[i] > v8 ? 1 : 0;
                    int pf = 0, pb = 7, ties_left = 8 - n_gt;
                    if (!(v8 > -INFINITY)) ties_left = 0;      // fewer than 8 real rows: nothing ties with the empty value
                    __syncwarp();
                    tmem_ld_wait();
                    tmem_ld32(t_base + 32, vb);
                    boot_emit32(va, allow[0], masked, 0, v8, pf, pb, ties_left, ent);
                    __syncwarp();
                    tmem_ld_wait();
                    tmem_ld32(t_base + 64, va);
                    boot_emit32(vb, allow[1], masked, 32, v8, pf, pb, ties_left, ent);
                    __syncwarp();
                    tmem_ld_wait();
                    tmem_ld32(t_base + 96, vb);
                    boot_emit32(va, allow[2], masked, 64, v8, pf, pb, ties_left, ent);
                    __syncwarp();
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (CG == 1 || leader) mbar_arrive(smem_u32(tempty_bar + acc));
                        else mbar_arrive_remote(smem_u32(tempty_bar + acc), 0);
                    }
                    boot_emit32(vb, allow[3], masked, 96, v8, pf, pb, ties_left, ent);
                    // slot block (group ordinal m, tile half) of the query's buffer: BOOT_J keys, the best one first, 0 = empty
                    if (q < plan.nq) {
                        u64 kk[8];
                        u64 best = 0ull;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint2 e = ent[i];
                            kk[i] = e.y != 0xFFFFFFFFu ? avs_make_key(__uint_as_float(e.x), (uint32_t)(row0 + (int)e.y)) : 0ull;
                            best = kk[i] > best ? kk[i] : best;
                        }
                        // the best key goes to slot 0 (the level select reads the slot-0 keys first): swap it with whatever sits there
                        const u64 first = kk[0];
#pragma unroll
                        for (int i = 1; i < 8; ++i) kk[i] = kk[i] == best ? first : kk[i];
                        kk[0] = best;
                        ulonglong2* dst = reinterpret_cast<ulonglong2*>(my_cand + ((size_t)m * 2 + half) * AVS_BOOT_J);
#pragma unroll
                        for (int i = 0; i < 8; i += 2) dst[i >> 1] = make_ulonglong2(kk[i], kk[i + 1]);
                    }
                } else {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (CG == 1 || leader) mbar_arrive(smem_u32(tempty_bar + acc));
                        else mbar_arrive_remote(smem_u32(tempty_bar + acc), 0);
                    }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                continue;
            }
            uint32_t va[32], vb[32];
            __syncwarp();
            tmem_ld32(t_base, va);
            tmem_ld_wait();
            tmem_ld32(t_base + 32, vb);                // in flight while the previous chunk is examined
            process(va, 0);
            __syncwarp();
            tmem_ld_wait();
            tmem_ld32(t_base + 64, va);
            process(vb, 32);
            __syncwarp();
            tmem_ld_wait();
            tmem_ld32(t_base + 96, vb);
            process(va, 64);
            __syncwarp();
            tmem_ld_wait();
            // the accumulator has been copied out completely: hand it back to the MMA before the last compare pass
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 1 || leader) mbar_arrive(smem_u32(tempty_bar + acc));
                else mbar_arrive_remote(smem_u32(tempty_bar + acc), 0);
            }
            process(vb, 96);
            if (n_raw) drain();
            if (n_stash) reserve();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        };
        if (lv.dense == 2 && plan.nq <= 8) run_tiles(std::integral_constant<int, 3>{});
        else if (lv.dense == 2) run_tiles(std::integral_constant<int, 2>{});
        else if (lv.dense == 1) run_tiles(std::integral_constant<int, 1>{});
        else run_tiles(std::integral_constant<int, 0>{});
        flush_pending();
        stamp(1 + 4 * l);                              // this CTA's tiles of the level are done
        if (plan.trace != nullptr && threadIdx.x == 128 && blockIdx.x < 256) {   // per-CTA finish time of the level (load balance)
            u64 tn; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tn));
            plan.trace[64 + l * 256 + blockIdx.x] = tn;
        }
        // ---- level done on this CTA: wait for the whole grid, then select (warp per query), then the next level ----
        grid_barrier_epi(plan.gbar, epoch, plan.err);
        stamp(2 + 4 * l);                              // the whole grid has finished the level
        {
            const bool final_level = plan.last_is_final && l == plan.n_levels - 1;
            u64* const list = reinterpret_cast<u64*>(warp_stash);             // 2 KB of the (now idle) survivor stash of this warp
            const int dense_total = lv.dense == 2 ? (int)(lv.n_visit * 2 * AVS_BOOT_J) : lv.dense ? (int)(lv.n_visit * AVS_GROUP_ROWS) : 0;
            if (plan.nq <= (int)gridDim.x) {           // at most one query per CTA: its 8 epilogue warps share the select
                if ((int)blockIdx.x < plan.nq) {
                    u64* const tr = (plan.trace != nullptr && blockIdx.x == 0) ? plan.trace + AVS_TRACE_SEL + l * 16 : nullptr;
                    if (!cta_select_fast(plan, (int)blockIdx.x, (int)threadIdx.x - 128, plan.j_rank[l], final_level, dense_total,
                                         plan.k_eps[l], stash_smem, []() { epi_sync(); }, tr))
                        cta_select_level(plan, (int)blockIdx.x, warp - 4, lane, plan.j_rank[l], final_level, dense_total, plan.k_eps[l],
                                         stash_smem, []() { epi_sync(); });
                }
            } else {
                for (int q = blockIdx.x * 8 + (warp - 4); q < plan.nq; q += gridDim.x * 8)
                    warp_select_level(plan, q, lane, plan.j_rank[l], final_level, dense_total, plan.k_eps[l], list, lv.dense == 2);
            }
        }
        stamp(3 + 4 * l);                              // this CTA's share of the selects is done
        if (l + 1 < plan.n_levels) grid_barrier_epi(plan.gbar, epoch, plan.err);
        stamp(4 + 4 * l);                              // every query has its new threshold
        }
    }

    tc_fence_before();
    if constexpr (CG == 1) __syncthreads(); else cluster_sync_all();
    if (warp == 2) {
        if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D bf16 row-major [rows, dpad] tensor, box = 64 elements (128 B) x box_rows, 128-byte swizzle
int make_map(CUtensorMap* map, const void* base, int64_t rows, int dpad, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { avs_set_error("cuTensorMapEncodeTiled is not available from the driver"); return AVS_E_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)dpad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)dpad * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { avs_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %lld, dpad %d)", (int)r, (long long)rows, dpad); return AVS_E_CUDA; }
    return AVS_OK;
}

template <int CG>
int launch(avs_store* s, int nq, const AvsScanPlan& plan, cudaStream_t st) {
    using C = Cfg<CG>;
    const int q_block = BLOCK_M * CG;
    const int nq_pad = (nq + 255) / 256 * 256;            // prep_queries pads to 256: whole blocks for both CG
    const int n_qblocks = (nq + q_block - 1) / q_block;
    CUtensorMap mq, mx;
    AVS_CHECK(make_map(&mq, s->sc.qb, nq_pad, s->dpad, BLOCK_M));
    AVS_CHECK(make_map(&mx, s->xb, s->capacity, s->dpad, C::LOAD_N));
    static bool attr[64] = {};   // function attributes are per device
    static int max_clusters[64] = {};
    if (!attr[s->device & 63]) {
        AVS_CUDA(cudaFuncSetAttribute(scan_gemm_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        // the grid barrier between levels needs every CTA resident at once: ask the runtime how many clusters that is
        int nc = s->num_sms / CG;
        if (CG > 1) {
            cudaLaunchConfig_t qc = {};
            qc.gridDim = dim3((unsigned)(s->num_sms / CG * CG));
            qc.blockDim = dim3(GEMM_THREADS);
            qc.dynamicSmemBytes = C::SMEM_BYTES;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = CG; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            qc.attrs = qa; qc.numAttrs = 1;
            int got = 0;
            if (cudaOccupancyMaxActiveClusters(&got, scan_gemm_kernel<CG>, &qc) == cudaSuccess && got > 0) nc = got < nc ? got : nc;
            else cudaGetLastError();
        } else {
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_gemm_kernel<CG>, GEMM_THREADS, C::SMEM_BYTES) == cudaSuccess && per_sm >= 1) nc = s->num_sms;
            else cudaGetLastError();
        }
        max_clusters[s->device & 63] = nc < 1 ? 1 : nc;
        attr[s->device & 63] = true;
    }
    int64_t most = 1;                                     // the busiest level decides how many clusters are useful
    for (int l = 0; l < plan.n_levels; ++l) { const int64_t t = plan.lv[l].n_visit * n_qblocks; most = t > most ? t : most; }
    int64_t clusters = max_clusters[s->device & 63];
    if (clusters > most) clusters = most;
    if (clusters < 1) clusters = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * CG));
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = s->opt_pdl ? 2 : 1;
    AVS_CUDA(cudaLaunchKernelEx(&cfg, scan_gemm_kernel<CG>, mq, mx, s->count, n_qblocks, s->dpad / BLOCK_K, plan,
                                (const uint32_t*)s->filter));
    s->st_launches++;
    return AVS_OK;
}

}  // namespace

// One launch: every level in `plan` scanned by the tensor-core kernel with the level selects fused in.
int avs_launch_scan_gemm(avs_store* s, int nq, const AvsScanPlan& plan, cudaStream_t st) {
    if (s->opt_cta_group == 1 || (nq <= BLOCK_M && s->opt_cta_group_small == 1)) return launch<1>(s, nq, plan, st);
    return launch<2>(s, nq, plan, st);
}

void avs_gemm_state_free(avs_store* s) { s->gemm_state = nullptr; }
