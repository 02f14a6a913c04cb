// Device-resident embedding store: fp32 master + bf16 scan copy built by the
// normalise-on-insert kernel (K1).  Replaces the storage side of
// `MilvusClient.create_collection / insert` (/root/reference/milvus/RAG.py:54-57,541-544).
#include <stdarg.h>
#include <stdio.h>

#include "avs_internal.h"

static thread_local char g_err[512] = "";

void avs_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* avs_last_error(void) { return g_err; }
extern "C" const char* avs_version(void) { return "avs 0.1 (sm_100a)"; }

// ---------------------------------------------------------------------------------------------
// K1: normalise-on-insert.  One warp per row.  Pass 1 reads the fp32 row with 128-bit loads and
// reduces the squared norm; pass 2 re-reads it (L1/L2 hit), scales, rounds to bf16 and stores
// 128-bit packets; the rounding residual ||bf16(x^) - x^|| feeds the store-wide r_max that the
// search certificate uses.  HBM-bound: 4 B read + 2 B written per element.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
    atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void __launch_bounds__(256) normalize_rows_kernel(const float* __restrict__ master,
                                                             __nv_bfloat16* __restrict__ xb,
                                                             float* __restrict__ inv_norm,
                                                             float* __restrict__ gstat, int64_t row0,
                                                             int64_t n, int dim, int dpad, int metric) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    float local_r = 0.f, local_n = 0.f;
    for (int64_t r = warp; r < n; r += nwarps) {
        const int64_t row = row0 + r;
        const float* x = master + row * dim;
        double ss = 0.0;  // float64: keeps 1/||x|| accurate to ~2 ulp so the certificate slack stays tight
        const bool vec = ((dim & 3) == 0) && ((((uintptr_t)x) & 15) == 0);
        if (vec) {
            const float4* x4 = reinterpret_cast<const float4*>(x);
            for (int c = lane; c < (dim >> 2); c += 32) {
                float4 v = __ldg(x4 + c);
                ss += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
            }
        } else {
            for (int c = lane; c < dim; c += 32) {
                float v = __ldg(x + c);
                ss += (double)v * v;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float nrm = (float)sqrt(ss);
        const float inv = ss > 0.0 ? (float)(1.0 / sqrt(ss)) : 0.f;
        const float scale = (metric == AVS_METRIC_COSINE) ? inv : 1.0f;
        float res = 0.f;
        __nv_bfloat16* o = xb + row * dpad;
        // 8 elements (16 bytes of bf16) per lane per step; dpad is a multiple of 64
        for (int c = lane * 8; c < dpad; c += 256) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (c + i < dim) ? __ldg(x + c + i) * scale : 0.f;
            uint4 pk;
            uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                __nv_bfloat16 a = __float2bfloat16_rn(v[2 * i]);
                __nv_bfloat16 b = __float2bfloat16_rn(v[2 * i + 1]);
                float da = __bfloat162float(a) - v[2 * i];
                float db = __bfloat162float(b) - v[2 * i + 1];
                res += da * da + db * db;
                pw[i] = (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
            }
            *reinterpret_cast<uint4*>(o + c) = pk;
        }
        res = warp_sum(res);
        if (lane == 0) {
            inv_norm[row] = inv;
            local_r = fmaxf(local_r, sqrtf(res));
            local_n = fmaxf(local_n, nrm);
        }
    }
    if (lane == 0) {
        if (local_r > 0.f) atomic_max_nonneg(gstat + 0, local_r);
        if (local_n > 0.f) atomic_max_nonneg(gstat + 1, local_n);
    }
}

// Single-pass variant for dim % 4 == 0 and dim <= 128 * MAXV: the row stays in registers between the norm and
// the rounding pass, so every fp32 element is read from HBM exactly once (4 B in, 2 B out per element).
template <int MAXV>
__global__ void __launch_bounds__(256) normalize_rows_reg_kernel(const float* __restrict__ master,
                                                                 __nv_bfloat16* __restrict__ xb,
                                                                 float* __restrict__ inv_norm,
                                                                 float* __restrict__ gstat, int64_t row0,
                                                                 int64_t n, int dim, int dpad, int metric) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int nvec = dim >> 2, nvec_pad = dpad >> 2;
    float local_r = 0.f, local_n = 0.f;
    for (int64_t r = warp; r < n; r += nwarps) {
        const int64_t row = row0 + r;
        const float4* x4 = reinterpret_cast<const float4*>(master + row * dim);
        float4 buf[MAXV];
        double ss = 0.0;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int c = lane + 32 * i;
            buf[i] = c < nvec ? __ldcs(x4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            ss += (double)buf[i].x * buf[i].x + (double)buf[i].y * buf[i].y + (double)buf[i].z * buf[i].z + (double)buf[i].w * buf[i].w;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float nrm = (float)sqrt(ss);
        const float inv = ss > 0.0 ? (float)(1.0 / sqrt(ss)) : 0.f;
        const float scale = (metric == AVS_METRIC_COSINE) ? inv : 1.0f;
        float res = 0.f;
        uint2* o = reinterpret_cast<uint2*>(xb + row * dpad);
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int c = lane + 32 * i;
            if (c < nvec_pad) {
                const float v0 = buf[i].x * scale, v1 = buf[i].y * scale, v2 = buf[i].z * scale, v3 = buf[i].w * scale;
                const __nv_bfloat16 b0 = __float2bfloat16_rn(v0), b1 = __float2bfloat16_rn(v1), b2 = __float2bfloat16_rn(v2), b3 = __float2bfloat16_rn(v3);
                const float d0 = __bfloat162float(b0) - v0, d1 = __bfloat162float(b1) - v1, d2 = __bfloat162float(b2) - v2, d3 = __bfloat162float(b3) - v3;
                res += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
                uint2 pk;
                pk.x = (uint32_t)__bfloat16_as_ushort(b0) | ((uint32_t)__bfloat16_as_ushort(b1) << 16);
                pk.y = (uint32_t)__bfloat16_as_ushort(b2) | ((uint32_t)__bfloat16_as_ushort(b3) << 16);
                o[c] = pk;
            }
        }
        res = warp_sum(res);
        if (lane == 0) {
            inv_norm[row] = inv;
            local_r = fmaxf(local_r, sqrtf(res));
            local_n = fmaxf(local_n, nrm);
        }
    }
    if (lane == 0) {
        if (local_r > 0.f) atomic_max_nonneg(gstat + 0, local_r);
        if (local_n > 0.f) atomic_max_nonneg(gstat + 1, local_n);
    }
}

// ---------------------------------------------------------------------------------------------
// Synthetic rows (benchmark utility).  Element (row, col) of stream `seed`:
//   z = mix(mix(seed + row*G1) ^ (col+1)*G2)         (splitmix64 finaliser)
//   v = sum of the four 16-bit fields of z - 131070   (Irwin-Hall(4), integer, symmetric)
//   x = float(double(v) / sqrt(double(sum_c v^2)))    (integer sum of squares: exact)
// Integer arithmetic + correctly rounded sqrt/div only, so the numpy replay in
// autostyle-tts_b200/synth.py is bit-identical.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ u64 avs_mix64(u64 z) {
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
__host__ __device__ __forceinline__ int avs_synth_int(u64 seed, u64 row, u64 col) {
    u64 z = avs_mix64(seed + row * 0x9E3779B97F4A7C15ull);
    z = avs_mix64(z ^ ((col + 1) * 0xD1B54A32D192ED03ull));
    int v = (int)(z & 0xFFFF) + (int)((z >> 16) & 0xFFFF) + (int)((z >> 32) & 0xFFFF) + (int)(z >> 48);
    return v - 131070;
}

__global__ void __launch_bounds__(256) synth_rows_kernel(float* __restrict__ master, int64_t* __restrict__ ids,
                                                         int64_t row0, int64_t n, int dim, u64 seed,
                                                         int64_t first_row, int64_t id_base) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < n; r += nwarps) {
        const u64 srow = (u64)(first_row + r);
        u64 ss = 0;
        for (int c = lane; c < dim; c += 32) {
            long long v = avs_synth_int(seed, srow, (u64)c);
            ss += (u64)(v * v);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const double nrm = sqrt((double)ss);
        float* x = master + (row0 + r) * dim;
        for (int c = lane; c < dim; c += 32) {
            int v = avs_synth_int(seed, srow, (u64)c);
            x[c] = nrm > 0.0 ? (float)((double)v / nrm) : 0.f;
        }
        if (lane == 0) ids[row0 + r] = id_base + first_row + r;
    }
}

__global__ void iota_ids_kernel(int64_t* ids, int64_t row0, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ids[row0 + i] = row0 + i;
}

// ---------------------------------------------------------------------------------------------
extern "C" int avs_set_filter(avs_store* s, const uint32_t* bitmap_host, int64_t n_bits);
static int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

static int alloc_arrays(avs_store* s, int64_t cap, float** master, __nv_bfloat16** xb, float** inv,
                        int64_t** ids) {
    size_t bm = (size_t)cap * s->dim * sizeof(float);
    size_t bx = (size_t)cap * s->dpad * sizeof(__nv_bfloat16);
    if (cudaMalloc(master, bm) != cudaSuccess || cudaMalloc(xb, bx) != cudaSuccess ||
        cudaMalloc(inv, (size_t)cap * sizeof(float)) != cudaSuccess ||
        cudaMalloc(ids, (size_t)cap * sizeof(int64_t)) != cudaSuccess) {
        cudaGetLastError();
        avs_set_error("out of device memory allocating %lld rows x %d dims", (long long)cap, s->dim);
        return AVS_E_NOMEM;
    }
    // padding rows of the bf16 copy must be zero: tiles are always whole groups
    AVS_CUDA(cudaMemset(*xb, 0, bx));
    AVS_CUDA(cudaMemset(*inv, 0, (size_t)cap * sizeof(float)));
    // the memsets run on the legacy default stream and are asynchronous to the host for device memory; the rows
    // that follow are written on the caller's stream, which may be non-blocking (every non-default torch stream
    // is): finish the clears before anything can be ordered against them
    AVS_CUDA(cudaDeviceSynchronize());
    return AVS_OK;
}

extern "C" int avs_create(int device, int dim, int metric, int64_t capacity, avs_store** out) {
    if (!out) { avs_set_error("avs_create: out is NULL"); return AVS_E_INVALID; }
    *out = nullptr;
    if (dim <= 0 || dim > 32768) { avs_set_error("avs_create: dim %d out of range [1, 32768]", dim); return AVS_E_INVALID; }
    if (metric != AVS_METRIC_COSINE && metric != AVS_METRIC_IP) { avs_set_error("avs_create: unknown metric %d", metric); return AVS_E_INVALID; }
    if (capacity < 0) { avs_set_error("avs_create: negative capacity"); return AVS_E_INVALID; }
    int ndev = 0;
    AVS_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { avs_set_error("avs_create: device %d not present (%d visible)", device, ndev); return AVS_E_INVALID; }
    AVS_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    AVS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { avs_set_error("avs_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return AVS_E_INVALID; }
    avs_store* s = new avs_store();
    s->device = device;
    s->dim = dim;
    s->dpad = (int)round_up(dim, 64);
    s->metric = metric;
    s->num_sms = prop.multiProcessorCount;
    s->capacity = round_up(capacity > 0 ? capacity : AVS_GROUP_ROWS, AVS_GROUP_ROWS);
    int rc = alloc_arrays(s, s->capacity, &s->master, &s->xb, &s->inv_norm, &s->ids);
    if (rc != AVS_OK) { avs_destroy(s); return rc; }
    if (cudaMalloc(&s->gstat, 4 * sizeof(float)) != cudaSuccess || cudaMalloc(&s->dstat, 8 * sizeof(u64)) != cudaSuccess) {
        avs_set_error("out of device memory (stats)");
        avs_destroy(s);
        return AVS_E_NOMEM;
    }
    cudaMemset(s->gstat, 0, 4 * sizeof(float));
    cudaMemset(s->dstat, 0, 8 * sizeof(u64));
    if (cudaHostAlloc((void**)&s->h_stats, 8 * sizeof(u64), cudaHostAllocDefault) == cudaSuccess) memset(s->h_stats, 0, 8 * sizeof(u64));
    else { cudaGetLastError(); s->h_stats = nullptr; }
    *out = s;
    return AVS_OK;
}

extern "C" int avs_destroy(avs_store* s) {
    if (!s) return AVS_OK;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    avs_comm_free(s);
    avs_gemm_state_free(s);
    avs_scratch_free(s);
    cudaFree(s->master);
    cudaFree(s->xb);
    cudaFree(s->inv_norm);
    cudaFree(s->ids);
    cudaFree(s->gstat);
    cudaFree(s->filter);
    cudaFree(s->dstat);
    if (s->h_stats) cudaFreeHost(s->h_stats);
    for (cudaEvent_t e : s->tev) cudaEventDestroy(e);
    cudaGetLastError();
    delete s;
    return AVS_OK;
}

extern "C" int avs_reserve(avs_store* s, int64_t capacity) {
    if (!s) { avs_set_error("avs_reserve: NULL store"); return AVS_E_INVALID; }
    if (capacity <= s->capacity) return AVS_OK;
    AVS_CUDA(cudaSetDevice(s->device));
    AVS_CUDA(cudaDeviceSynchronize());
    int64_t cap = round_up(capacity, AVS_GROUP_ROWS);
    float* m = nullptr; __nv_bfloat16* x = nullptr; float* inv = nullptr; int64_t* ids = nullptr;
    int rc = alloc_arrays(s, cap, &m, &x, &inv, &ids);
    if (rc != AVS_OK) { cudaFree(m); cudaFree(x); cudaFree(inv); cudaFree(ids); return rc; }
    AVS_CUDA(cudaMemcpy(m, s->master, (size_t)s->count * s->dim * sizeof(float), cudaMemcpyDeviceToDevice));
    AVS_CUDA(cudaMemcpy(x, s->xb, (size_t)s->count * s->dpad * sizeof(__nv_bfloat16), cudaMemcpyDeviceToDevice));
    AVS_CUDA(cudaMemcpy(inv, s->inv_norm, (size_t)s->count * sizeof(float), cudaMemcpyDeviceToDevice));
    AVS_CUDA(cudaMemcpy(ids, s->ids, (size_t)s->count * sizeof(int64_t), cudaMemcpyDeviceToDevice));
    AVS_CUDA(cudaDeviceSynchronize());   // device-to-device copies return before they ran (see alloc_arrays)
    cudaFree(s->master); cudaFree(s->xb); cudaFree(s->inv_norm); cudaFree(s->ids);
    s->master = m; s->xb = x; s->inv_norm = inv; s->ids = ids;
    s->capacity = cap;
    avs_gemm_state_free(s);  // tensor maps point at the old allocation
    return AVS_OK;
}

static int launch_normalize(avs_store* s, int64_t row0, int64_t n, cudaStream_t st) {
    if (n <= 0) return AVS_OK;
    int64_t blocks = (n + 7) / 8;
    int64_t maxb = (int64_t)s->num_sms * 8;
    if (blocks > maxb) blocks = maxb;
    const bool reg_path = (s->dim % 4 == 0) && (s->dpad / 4 <= 32 * 16);   // master rows are 16-byte aligned then
    if (reg_path && s->dpad / 4 <= 32 * 8)
        normalize_rows_reg_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(s->master, s->xb, s->inv_norm, s->gstat, row0, n,
                                                                       s->dim, s->dpad, s->metric);
    else if (reg_path)
        normalize_rows_reg_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(s->master, s->xb, s->inv_norm, s->gstat, row0, n,
                                                                        s->dim, s->dpad, s->metric);
    else
        normalize_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(s->master, s->xb, s->inv_norm, s->gstat, row0, n,
                                                                s->dim, s->dpad, s->metric);
    s->st_launches++;
    AVS_CUDA(cudaGetLastError());
    return AVS_OK;
}

static bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

extern "C" int avs_insert(avs_store* s, const float* rows, const int64_t* ids, int64_t n, void* stream) {
    if (!s) { avs_set_error("avs_insert: NULL store"); return AVS_E_INVALID; }
    if (n < 0 || (n > 0 && !rows)) { avs_set_error("avs_insert: bad rows/n"); return AVS_E_INVALID; }
    if (n == 0) return AVS_OK;
    if (s->count + n > 0xFFFFFFF0ll) { avs_set_error("avs_insert: more than 2^32 rows per store"); return AVS_E_NOMEM; }
    AVS_CUDA(cudaSetDevice(s->device));
    if (s->filter) AVS_CHECK(avs_set_filter(s, nullptr, 0));   // a row bitmap is only valid for the rows it was built on
    cudaStream_t st = (cudaStream_t)stream;
    if (s->count + n > s->capacity) {
        int64_t want = s->capacity * 2 > s->count + n ? s->capacity * 2 : s->count + n;
        AVS_CHECK(avs_reserve(s, want));
    }
    const int64_t row0 = s->count;
    AVS_CUDA(cudaMemcpyAsync(s->master + row0 * s->dim, rows, (size_t)n * s->dim * sizeof(float),
                             is_device_ptr(rows) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (ids) {
        AVS_CUDA(cudaMemcpyAsync(s->ids + row0, ids, (size_t)n * sizeof(int64_t),
                                 is_device_ptr(ids) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    } else {
        iota_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s->ids, row0, n);
        s->st_launches++;
        AVS_CUDA(cudaGetLastError());
    }
    AVS_CHECK(launch_normalize(s, row0, n, st));
    s->count += n;
    return AVS_OK;
}

extern "C" int avs_fill_synthetic(avs_store* s, uint64_t seed, int64_t first_row, int64_t n, int64_t id_base,
                                  void* stream) {
    if (!s || n < 0 || first_row < 0) { avs_set_error("avs_fill_synthetic: bad arguments"); return AVS_E_INVALID; }
    if (n == 0) return AVS_OK;
    if (s->count + n > 0xFFFFFFF0ll) { avs_set_error("avs_fill_synthetic: more than 2^32 rows per store"); return AVS_E_NOMEM; }
    AVS_CUDA(cudaSetDevice(s->device));
    if (s->filter) AVS_CHECK(avs_set_filter(s, nullptr, 0));
    cudaStream_t st = (cudaStream_t)stream;
    if (s->count + n > s->capacity) AVS_CHECK(avs_reserve(s, s->count + n));
    const int64_t row0 = s->count;
    int64_t blocks = (n + 7) / 8;
    int64_t maxb = (int64_t)s->num_sms * 8;
    if (blocks > maxb) blocks = maxb;
    synth_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(s->master, s->ids, row0, n, s->dim, (u64)seed, first_row,
                                                        id_base);
    s->st_launches++;
    AVS_CUDA(cudaGetLastError());
    AVS_CHECK(launch_normalize(s, row0, n, st));
    s->count += n;
    return AVS_OK;
}

extern "C" int avs_set_filter(avs_store* s, const uint32_t* bitmap_host, int64_t n_bits) {
    if (!s) { avs_set_error("avs_set_filter: NULL store"); return AVS_E_INVALID; }
    AVS_CUDA(cudaSetDevice(s->device));
    if (!bitmap_host) {                               // clear: every row may be returned again
        s->filter_allowed = 0;
        if (s->filter) { AVS_CUDA(cudaDeviceSynchronize()); cudaFree(s->filter); s->filter = nullptr; s->filter_words = 0; }
        return AVS_OK;
    }
    if (n_bits != s->count) { avs_set_error("avs_set_filter: bitmap has %lld bits, the store %lld rows", (long long)n_bits, (long long)s->count); return AVS_E_INVALID; }
    const size_t words = (size_t)((n_bits + 31) / 32);
    // the tensor-core scan reads one bitmap word per 32-row chunk of WHOLE 256-row groups (padding chunks of the last
    // group included, masked afterwards): the allocation covers whole groups and the tail words are zero
    const size_t words_alloc = (size_t)(round_up(n_bits > 0 ? n_bits : 1, AVS_GROUP_ROWS) / 32);
    AVS_CUDA(cudaDeviceSynchronize());
    if (words_alloc > s->filter_words || !s->filter) {
        cudaFree(s->filter);
        s->filter = nullptr;
        if (cudaMalloc((void**)&s->filter, words_alloc * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); s->filter_words = 0; avs_set_error("out of device memory for the filter bitmap"); return AVS_E_NOMEM; }
        s->filter_words = words_alloc;
    }
    AVS_CUDA(cudaMemset(s->filter, 0, s->filter_words * sizeof(uint32_t)));
    int64_t allowed = 0;
    for (size_t w = 0; w < words; ++w) {
        uint32_t v = bitmap_host[w];
        if (w == words - 1 && (n_bits & 31)) v &= (1u << (n_bits & 31)) - 1;
        allowed += __builtin_popcount(v);
    }
    AVS_CUDA(cudaMemcpy(s->filter, bitmap_host, words * sizeof(uint32_t), cudaMemcpyHostToDevice));
    AVS_CUDA(cudaDeviceSynchronize());
    s->filter_allowed = allowed;
    return AVS_OK;
}

extern "C" int64_t avs_count(const avs_store* s) { return s ? s->count : 0; }
extern "C" int avs_dim(const avs_store* s) { return s ? s->dim : 0; }
extern "C" int avs_metric(const avs_store* s) { return s ? s->metric : 0; }

extern "C" int avs_get_rows(avs_store* s, int64_t first, int64_t n, float* out, void* stream) {
    if (!s || !out || first < 0 || n < 0 || first + n > s->count) { avs_set_error("avs_get_rows: range [%lld, %lld) outside [0, %lld)", (long long)first, (long long)(first + n), (long long)(s ? s->count : 0)); return AVS_E_INVALID; }
    AVS_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    bool dev = is_device_ptr(out);
    AVS_CUDA(cudaMemcpyAsync(out, s->master + first * s->dim, (size_t)n * s->dim * sizeof(float),
                             dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (!dev) AVS_CUDA(cudaStreamSynchronize(st));
    return AVS_OK;
}

extern "C" int avs_get_ids(avs_store* s, int64_t first, int64_t n, int64_t* out, void* stream) {
    if (!s || !out || first < 0 || n < 0 || first + n > s->count) { avs_set_error("avs_get_ids: range outside the store"); return AVS_E_INVALID; }
    AVS_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    bool dev = is_device_ptr(out);
    AVS_CUDA(cudaMemcpyAsync(out, s->ids + first, (size_t)n * sizeof(int64_t),
                             dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (!dev) AVS_CUDA(cudaStreamSynchronize(st));
    return AVS_OK;
}

// ---------------------------------------------------------------------------------------------
// Bulk snapshot of a device store.  The reference's persistence is the Milvus Lite SQLite
// file: one protobuf blob per row (/root/reference/milvus/RAG.py:541-544 writes 130 of them), which the drop-in keeps
// for small collections (milvus_lite_db.py).  A 10^7-10^8-row store is saved raw instead:
//   header (64 bytes) | ids [count] int64 | master rows [count, dim] fp32 exactly as inserted
// The bf16 scan copy and the inverse norms are NOT stored: the normalise-on-insert kernel rebuilds them from the master
// at HBM speed, far faster than reading another 50 % from disk, and bit-identically to the original insert.
// Copies run through two pinned staging buffers so the PCIe transfer of one chunk overlaps the file I/O of the other.
// ---------------------------------------------------------------------------------------------
struct AvsSnapshotHeader {
    char magic[8];            // "AVSSNAP1"
    int32_t dim, metric;
    int64_t count;
    int64_t reserved[5];
};
static_assert(sizeof(AvsSnapshotHeader) == 64, "snapshot header is 64 bytes");
#define AVS_SNAP_CHUNK ((size_t)64 << 20)

static int snap_io_device(FILE* f, void* dev, size_t bytes, bool to_file, char* pin[2], cudaStream_t st[2], cudaEvent_t ev[2]) {
    size_t off = 0;
    int slot = 0;
    size_t pending[2] = {0, 0}, pending_off[2] = {0, 0};
    while (off < bytes || pending[0] || pending[1]) {
        if (pending[slot]) {                       // finish what this slot started two chunks ago
            AVS_CUDA(cudaEventSynchronize(ev[slot]));
            if (to_file && fwrite(pin[slot], 1, pending[slot], f) != pending[slot]) { avs_set_error("snapshot: short write"); return AVS_E_INVALID; }
            pending[slot] = 0;
        }
        if (off < bytes) {
            const size_t n = bytes - off < AVS_SNAP_CHUNK ? bytes - off : AVS_SNAP_CHUNK;
            if (to_file) {
                AVS_CUDA(cudaMemcpyAsync(pin[slot], (char*)dev + off, n, cudaMemcpyDeviceToHost, st[slot]));
            } else {
                if (fread(pin[slot], 1, n, f) != n) { avs_set_error("snapshot: file is truncated"); return AVS_E_INVALID; }
                AVS_CUDA(cudaMemcpyAsync((char*)dev + off, pin[slot], n, cudaMemcpyHostToDevice, st[slot]));
            }
            AVS_CUDA(cudaEventRecord(ev[slot], st[slot]));
            pending[slot] = n; pending_off[slot] = off;
            off += n;
        }
        slot ^= 1;
    }
    (void)pending_off;
    return AVS_OK;
}

struct SnapIo {
    char* pin[2] = {nullptr, nullptr};
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    FILE* f = nullptr;
    int open(const char* path, const char* mode) {
        f = fopen(path, mode);
        if (!f) { avs_set_error("snapshot: cannot open %s", path); return AVS_E_INVALID; }
        return AVS_OK;
    }
    int device_init() {
        for (int i = 0; i < 2; ++i) {
            if (cudaHostAlloc((void**)&pin[i], AVS_SNAP_CHUNK, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); avs_set_error("snapshot: out of pinned host memory"); return AVS_E_NOMEM; }
            AVS_CUDA(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
            AVS_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
        return AVS_OK;
    }
    ~SnapIo() {
        for (int i = 0; i < 2; ++i) {
            if (st[i]) { cudaStreamSynchronize(st[i]); cudaStreamDestroy(st[i]); }
            if (ev[i]) cudaEventDestroy(ev[i]);
            if (pin[i]) cudaFreeHost(pin[i]);
        }
        if (f) fclose(f);
        cudaGetLastError();
    }
};

extern "C" int avs_save(avs_store* s, const char* path) {
    if (!s || !path) { avs_set_error("avs_save: NULL argument"); return AVS_E_INVALID; }
    SnapIo io;
    AVS_CHECK(io.open(path, "wb"));
    AVS_CUDA(cudaSetDevice(s->device));
    AVS_CUDA(cudaDeviceSynchronize());
    AVS_CHECK(io.device_init());
    AvsSnapshotHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "AVSSNAP1", 8);
    h.dim = s->dim; h.metric = s->metric; h.count = s->count;
    if (fwrite(&h, sizeof(h), 1, io.f) != 1) { avs_set_error("snapshot: short write"); return AVS_E_INVALID; }
    AVS_CHECK(snap_io_device(io.f, s->ids, (size_t)s->count * sizeof(int64_t), true, io.pin, io.st, io.ev));
    AVS_CHECK(snap_io_device(io.f, s->master, (size_t)s->count * s->dim * sizeof(float), true, io.pin, io.st, io.ev));
    if (fflush(io.f) != 0) { avs_set_error("snapshot: flush failed"); return AVS_E_INVALID; }
    return AVS_OK;
}

extern "C" int avs_load(const char* path, int device, avs_store** out) {
    if (!path || !out) { avs_set_error("avs_load: NULL argument"); return AVS_E_INVALID; }
    *out = nullptr;
    SnapIo io;
    AVS_CHECK(io.open(path, "rb"));
    AvsSnapshotHeader h;
    if (fread(&h, sizeof(h), 1, io.f) != 1 || memcmp(h.magic, "AVSSNAP1", 8) != 0) { avs_set_error("avs_load: %s is not a store snapshot", path); return AVS_E_INVALID; }
    if (h.dim <= 0 || h.dim > 32768 || h.count < 0 || h.count > 0xFFFFFFF0ll || (h.metric != AVS_METRIC_COSINE && h.metric != AVS_METRIC_IP)) {
        avs_set_error("avs_load: corrupt snapshot header (dim %d, metric %d, count %lld)", h.dim, h.metric, (long long)h.count);
        return AVS_E_INVALID;
    }
    avs_store* s = nullptr;
    AVS_CHECK(avs_create(device, h.dim, h.metric, h.count, &s));
    int rc = io.device_init();
    if (rc == AVS_OK) rc = snap_io_device(io.f, s->ids, (size_t)h.count * sizeof(int64_t), false, io.pin, io.st, io.ev);
    if (rc == AVS_OK) rc = snap_io_device(io.f, s->master, (size_t)h.count * h.dim * sizeof(float), false, io.pin, io.st, io.ev);
    if (rc == AVS_OK && cudaDeviceSynchronize() != cudaSuccess) { avs_set_error("avs_load: device copy failed"); rc = AVS_E_CUDA; }
    if (rc == AVS_OK) rc = launch_normalize(s, 0, h.count, 0);       // rebuilds the bf16 copy, the norms and the certificate's r_max
    if (rc == AVS_OK && cudaDeviceSynchronize() != cudaSuccess) { avs_set_error("avs_load: normalise failed"); rc = AVS_E_CUDA; }
    if (rc != AVS_OK) { avs_destroy(s); return rc; }
    s->count = h.count;
    *out = s;
    return AVS_OK;
}
