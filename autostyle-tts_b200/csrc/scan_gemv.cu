// K2: HBM-bound scan of the bf16 copy for small query batches (<= 8 queries per pass) with the
// top-k selection fused into the epilogue: no score matrix is written, a row is appended to the
// per-query candidate buffer only when its key beats the running threshold key.
//
// Layout: work unit = 4 consecutive rows; the units of the level's row groups are dealt round-robin to
// all warps of the grid, so neighbouring warps stream neighbouring packets.  A row is D_pad bf16 = D_pad/8 16-byte packets; lane l loads packets l, l+32, ... with
// ld.global.nc.L1::no_allocate.v4 (every warp-level load covers 512 contiguous bytes).  Queries
// stay in shared memory as fp32 (not rounded to bf16, which halves the certificate slack), split
// into two float4 planes so that the per-packet reads are bank-conflict free.
// Algorithmic bytes per launch: rows_visited * D_pad * 2.
#include "avs_internal.h"

#define GEMV_ROWS 4
#define GEMV_THREADS 256

__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

__device__ __forceinline__ float dot8(const uint4& v, const float4& a, const float4& b, float acc) {
    acc = fmaf(bf_lo(v.x), a.x, acc);
    acc = fmaf(bf_hi(v.x), a.y, acc);
    acc = fmaf(bf_lo(v.y), a.z, acc);
    acc = fmaf(bf_hi(v.y), a.w, acc);
    acc = fmaf(bf_lo(v.z), b.x, acc);
    acc = fmaf(bf_hi(v.z), b.y, acc);
    acc = fmaf(bf_lo(v.w), b.z, acc);
    acc = fmaf(bf_hi(v.w), b.w, acc);
    return acc;
}

// Reduce NV per-lane partial sums across the warp so that lane l ends with the total of value
// index idx(l) (transpose-reduce: NV-1 + log2(32/NV) shuffles instead of 5*NV).
template <int NV>
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[NV], int lane, int& idx) {
    idx = 0;
    int off = 16;
#pragma unroll
    for (int n = NV; n > 1; n >>= 1, off >>= 1) {
        const int h = n >> 1;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            float send = up ? v[i] : v[i + h];
            float keep = up ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
        if (up) idx += h;
    }
    float r = v[0];
    for (; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
    return r;
}

template <int NQ>
__global__ void __launch_bounds__(GEMV_THREADS, 2)
scan_gemv_kernel(const uint4* __restrict__ xb, int packets_per_row, int64_t n_rows,
                 const float* __restrict__ qf, int dpad, const u64* __restrict__ tau, AvsLevel lv,
                 u64* __restrict__ cand, int* __restrict__ cnt, int cap, const uint32_t* __restrict__ filt,
                 u64* __restrict__ dense, int dense_cap) {
    extern __shared__ float4 sq[];  // [NQ][2][packets_per_row]: plane 0 = first 4 floats of a packet
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < NQ * packets_per_row; i += GEMV_THREADS) {
        const int q = i / packets_per_row, c = i - q * packets_per_row;
        const float4* src = reinterpret_cast<const float4*>(qf + (size_t)q * dpad) + 2 * c;
        sq[(q * 2 + 0) * packets_per_row + c] = src[0];
        sq[(q * 2 + 1) * packets_per_row + c] = src[1];
    }
    __shared__ u64 tau_k[NQ];    // indexed by a runtime query slot after the transpose-reduce
    __shared__ float tau_f[NQ];
    if (threadIdx.x < NQ) {
        const u64 t = tau[threadIdx.x];
        tau_k[threadIdx.x] = t;
        tau_f[threadIdx.x] = t == 0 ? -INFINITY : avs_key_score(t);
    }
    __syncthreads();

    // work unit = 4 consecutive rows; units of the visited groups are dealt round-robin to all warps of the
    // grid, so neighbouring warps stream neighbouring 4-row packets and even the sparsest level fills the chip
    constexpr int UNITS = AVS_GROUP_ROWS / GEMV_ROWS;
    const int64_t n_units = lv.n_visit * UNITS;
    const int64_t gwarp = (int64_t)blockIdx.x * (GEMV_THREADS / 32) + warp;
    const int64_t n_gwarps = (int64_t)gridDim.x * (GEMV_THREADS / 32);
    for (int64_t u = gwarp; u < n_units; u += n_gwarps) {
        {
            const int64_t m = u / UNITS;
            const int64_t rbase = avs_level_group(lv, m) * AVS_GROUP_ROWS + (u - m * UNITS) * GEMV_ROWS;
            if (rbase >= n_rows) continue;
            float acc[GEMV_ROWS * NQ];
#pragma unroll
            for (int i = 0; i < GEMV_ROWS * NQ; ++i) acc[i] = 0.f;
            // rows past n_rows inside an allocated group are zero-filled, so reading them is safe
            const uint4* rp = xb + rbase * packets_per_row;
#pragma unroll 2
            for (int c = lane; c < packets_per_row; c += 32) {
                uint4 v[GEMV_ROWS];
#pragma unroll
                for (int r = 0; r < GEMV_ROWS; ++r) v[r] = ld_stream_u4(rp + (size_t)r * packets_per_row + c);
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const float4 a = sq[(q * 2 + 0) * packets_per_row + c];
                    const float4 b = sq[(q * 2 + 1) * packets_per_row + c];
#pragma unroll
                    for (int r = 0; r < GEMV_ROWS; ++r) acc[r * NQ + q] = dot8(v[r], a, b, acc[r * NQ + q]);
                }
            }
            int idx;
            const float s = warp_transpose_reduce<GEMV_ROWS * NQ>(acc, lane, idx);
            constexpr int DUP = 32 / (GEMV_ROWS * NQ);  // lanes holding the same total
            const bool owner = (DUP <= 1) || ((lane & (DUP - 1)) == 0);
            const int r = idx / NQ, q = idx - r * NQ;
            const int64_t row = rbase + r;
            if (lv.dense) {
                // threshold-free level: every visited row's key goes to its own slot (0 = padding / filtered out)
                if (owner) {
                    const bool keep = row < n_rows && (!filt || ((filt[row >> 5] >> (row & 31)) & 1u));
                    const int64_t slot = m * AVS_GROUP_ROWS + (u - m * UNITS) * GEMV_ROWS + r;
                    dense[(size_t)q * dense_cap + slot] = keep ? avs_make_key(s, (uint32_t)row) : 0ull;
                }
            } else if (owner && row < n_rows && s >= tau_f[q] && (!filt || ((filt[row >> 5] >> (row & 31)) & 1u))) {
                const u64 key = avs_make_key(s, (uint32_t)row);
                if (key >= tau_k[q]) {
                    const int pos = atomicAdd(cnt + q, 1);
                    if (pos < cap) cand[(size_t)q * cap + pos] = key;
                }
            }
        }
    }
}

template <int NQ>
static int launch_gemv(avs_store* s, int q0, const AvsLevel& lv, int cap, cudaStream_t st) {
    const int ppr = s->dpad / 8;
    const size_t smem = (size_t)NQ * 2 * ppr * sizeof(float4);
    static bool attr_set[64] = {};   // function attributes are per device
    if (!attr_set[s->device & 63]) {
        AVS_CUDA(cudaFuncSetAttribute(scan_gemv_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[s->device & 63] = true;
    }
    if (smem > 200 * 1024) { avs_set_error("gemv scan: %d queries x %d dims do not fit shared memory", NQ, s->dpad); return AVS_E_INVALID; }
    const int64_t want = (lv.n_visit * (AVS_GROUP_ROWS / GEMV_ROWS) + GEMV_THREADS / 32 - 1) / (GEMV_THREADS / 32);
    int64_t grid = want < (int64_t)s->num_sms * 2 ? want : (int64_t)s->num_sms * 2;
    if (grid < 1) grid = 1;
    scan_gemv_kernel<NQ><<<(unsigned)grid, GEMV_THREADS, smem, st>>>(
        reinterpret_cast<const uint4*>(s->xb), ppr, s->count, s->sc.qf + (size_t)q0 * s->dpad, s->dpad,
        s->sc.tau + q0, lv, s->sc.cand + (size_t)q0 * cap, s->sc.cnt + q0, cap, s->filter,
        s->sc.dense_buf ? s->sc.dense_buf + (size_t)q0 * AVS_DENSE_CAP : nullptr, AVS_DENSE_CAP);
    s->st_launches++;
    AVS_CUDA(cudaGetLastError());
    return AVS_OK;
}

// tau/cand/cnt are indexed from q0; idx inside the kernel is relative to the pass.
int avs_launch_scan_gemv(avs_store* s, int q0, int nq, const AvsLevel& lv, int cap, cudaStream_t st) {
    // a pass handles 1, 2, 4 or 8 queries; odd remainders use the next size up on padded (never
    // accepting) query slots, which prep_queries provides up to nq_pad.
    int width = nq <= 1 ? 1 : nq <= 2 ? 2 : nq <= 4 ? 4 : 8;
    // keep the fp32 query planes within shared memory: halve the pass width for very large dims
    while (width > 1 && (size_t)width * s->dpad * 4 > 96 * 1024) width >>= 1;
    for (int off = 0; off < nq; off += width) {
        int rc;
        switch (width) {
            case 1: rc = launch_gemv<1>(s, q0 + off, lv, cap, st); break;
            case 2: rc = launch_gemv<2>(s, q0 + off, lv, cap, st); break;
            case 4: rc = launch_gemv<4>(s, q0 + off, lv, cap, st); break;
            default: rc = launch_gemv<8>(s, q0 + off, lv, cap, st); break;
        }
        if (rc != AVS_OK) return rc;
    }
    return AVS_OK;
}
