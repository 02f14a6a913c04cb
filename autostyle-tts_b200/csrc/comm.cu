// Multi-GPU: row-sharded store, one process per GPU.  Every rank runs the
// local exact top-k on its shard, then ONE ncclAllGather of [nq, k] (float64 score, int64 id)
// pairs crosses NVLink and every rank merges world*k -> k with the same (score desc, id asc)
// comparator (K6).  NCCL is resolved with dlopen so that libavs.so loads without it and shares
// whichever libnccl.so.2 the process already holds (torch bundles one).
#include <dlfcn.h>
#include <nccl.h>

#include "avs_internal.h"

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.ok) return AVS_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) { avs_set_error("NCCL not found: dlopen(libnccl.so.2) failed: %s", dlerror()); return AVS_E_NCCL; }
#define LOAD(sym)                                                                              \
    g_nccl.sym = reinterpret_cast<decltype(g_nccl.sym)>(dlsym(g_nccl.lib, "nccl" #sym));       \
    if (!g_nccl.sym) { avs_set_error("NCCL symbol nccl" #sym " missing"); return AVS_E_NCCL; }
    LOAD(GetUniqueId)
    LOAD(CommInitRank)
    LOAD(CommDestroy)
    LOAD(AllGather)
    LOAD(GetErrorString)
#undef LOAD
    g_nccl.ok = true;
    return AVS_OK;
}
}  // namespace

#define AVS_NCCL(expr)                                                                         \
    do {                                                                                       \
        ncclResult_t _r = (expr);                                                              \
        if (_r != ncclSuccess) {                                                               \
            avs_set_error("%s failed: %s", #expr, g_nccl.GetErrorString(_r));                  \
            return AVS_E_NCCL;                                                                 \
        }                                                                                      \
    } while (0)

struct GatherItem { double s; int64_t id; };

__global__ void pack_gather_kernel(const double* __restrict__ s64, const int64_t* __restrict__ ids, GatherItem* out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { out[i].s = s64[i]; out[i].id = ids[i]; }
}

__device__ __forceinline__ bool item_better(const GatherItem& a, const GatherItem& b) {
    if (a.s != b.s) return a.s > b.s;
    return a.id < b.id;
}

// K6: one CTA per query; world*k <= 8*256 items sorted in shared memory.
__global__ void __launch_bounds__(256) shard_merge_kernel(const GatherItem* __restrict__ recv, int world, int nq, int k,
                                                          int64_t* __restrict__ out_ids, float* __restrict__ out_scores) {
    extern __shared__ unsigned char raw[];
    GatherItem* sm = reinterpret_cast<GatherItem*>(raw);
    const int q = blockIdx.x;
    const int total = world * k;
    int P = 32;
    while (P < total) P <<= 1;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        GatherItem it;
        if (i < total) {
            const int r = i / k, j = i - r * k;
            it = recv[((size_t)r * nq + q) * k + j];
            if (it.id == -1 && it.s == -INFINITY) it.id = INT64_MAX;  // padding sorts last
        } else { it.s = -INFINITY; it.id = INT64_MAX; }
        sm[i] = it;
    }
    __syncthreads();
    for (int k2 = 2; k2 <= P; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool desc = (i & k2) == 0;
                    const GatherItem a = sm[i], b = sm[ixj];
                    if (desc ? item_better(b, a) : item_better(a, b)) { sm[i] = b; sm[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const GatherItem it = sm[t];
        const bool pad = it.id == INT64_MAX && it.s == -INFINITY;
        out_ids[(size_t)q * k + t] = pad ? -1 : it.id;
        out_scores[(size_t)q * k + t] = pad ? -INFINITY : (float)it.s;
    }
}

// ---------------------------------------------------------------------------------------------
// Fused exchange + merge over NVLink peer memory (replaces pack -> ncclAllGather -> merge when connected).
// Every rank exposes one exchange region through CUDA IPC.  One kernel per search; its grid is never larger than
// what the device keeps resident at once, so no CTA ever waits for a CTA that has not been scheduled:
//   phase 1 (all of this CTA's queries, grid-stride): store this rank's k (score, id) items of the query straight
//            into slot [rank] of EVERY peer's region (remote stores over NVLink) and of its own,
//            __threadfence_system, publish flag[rank][q] = seq in every region;
//   phase 2 (same queries): spin until the local region shows seq for the query from every rank, merge world*k
//            items by (score desc, id asc), write the global top-k.
// Every CTA of every rank finishes phase 1 without waiting for anybody, hence every wait of phase 2 completes
// whatever the dispatch order.  Two parities alternate between searches: a rank can only be one search ahead of
// its slowest peer (it needs that peer's push to finish its own merge), so a region is never overwritten while it
// is still being read.  A wait that gives up (wall-clock bound, a peer died) poisons the query's output (id -1, score -inf) and
// bumps `timeouts`, which the host mirrors into pinned memory: the next call fails with AVS_E_NCCL.
// ---------------------------------------------------------------------------------------------
#define P2P_MAX_WORLD 8
#define P2P_ITEMS_PER_SRC (1 << 18)          // (score, id) items per source rank per parity (4 MiB)
#define P2P_MAX_NQ (1 << 14)
#define P2P_Q_FLOATS (1 << 23)               // query all-gather buffer per parity (32 MiB of fp32)
#define P2P_Q_CTAS 128                       // CTAs of the query all-gather kernel (one completion flag each)
#define P2P_TIMEOUT_MS_DEFAULT 120000        // wall-clock bound of a wait for a peer (option "p2p_timeout_ms")

__device__ __forceinline__ unsigned long long p2p_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Wait until *f shows seq.  The bound is WALL-CLOCK time (globaltimer), not a spin count: a peer that is merely late
// (its host is busy between two searches) must not be mistaken for a dead one.  Returns false when the wait gave up.
__device__ __forceinline__ bool p2p_wait_flag(volatile unsigned int* f, unsigned int seq, unsigned long long timeout_ns) {
    if (*f == seq) return true;
    const unsigned long long t0 = p2p_now_ns();
    for (;;) {
        for (int i = 0; i < 256; ++i)
            if (*f == seq) return true;
        if (p2p_now_ns() - t0 > timeout_ns) return false;
        __nanosleep(200);
    }
}

struct P2PRegion {                            // layout of one rank's exchange region
    GatherItem items[2][P2P_MAX_WORLD][P2P_ITEMS_PER_SRC];
    unsigned int flags[2][P2P_MAX_WORLD][P2P_MAX_NQ];
    float qbuf[2][P2P_Q_FLOATS];              // the query batch, every rank's slice pushed by its owner
    unsigned int qflags[2][P2P_MAX_WORLD][P2P_Q_CTAS];
    unsigned int timeouts;
};
struct P2PPeers { P2PRegion* r[P2P_MAX_WORLD]; };

struct P2PState {
    P2PRegion* local = nullptr;
    P2PPeers peers{};
    bool connected = false;
    unsigned int seq = 0, qseq = 0;
    unsigned int* h_timeouts = nullptr;       // pinned mirror of local->timeouts, refreshed after every exchange
    unsigned int seen_timeouts = 0;
    int max_ctas = 0;                         // co-resident CTAs of the exchange kernel on this device
    std::vector<cudaEvent_t> tev;             // event pairs around the exchange kernel (timing hook)
    size_t tev_used = 0;
};

__global__ void __launch_bounds__(256) p2p_exchange_merge_kernel(P2PPeers peers, int rank, int world, unsigned int seq,
                                                                 unsigned long long timeout_ns, int nq, int k, const double* __restrict__ s64,
                                                                 int64_t* __restrict__ out_ids, float* __restrict__ out_scores) {
    extern __shared__ unsigned char raw[];
    GatherItem* sm = reinterpret_cast<GatherItem*>(raw);
    __shared__ int s_dead;
    const int par = seq & 1;
    // phase 1: push my items of every query this CTA owns into every region (mine included), then publish
    for (int q = blockIdx.x; q < nq; q += gridDim.x) {
        const size_t slot = (size_t)q * k;
        for (int i = threadIdx.x; i < world * k; i += blockDim.x) {
            const int r = i / k, t = i - r * k;
            GatherItem it;
            it.s = s64[slot + t];
            it.id = out_ids[slot + t];
            peers.r[r]->items[par][rank][slot + t] = it;
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < world) {
            volatile unsigned int* f = &peers.r[threadIdx.x]->flags[par][rank][q];
            *f = seq;
        }
    }
    // phase 2: wait for every rank's items of my queries in MY region (wall-clock bounded wait: a missing peer must not hang the GPU)
    P2PRegion* mine = peers.r[rank];
    for (int q = blockIdx.x; q < nq; q += gridDim.x) {
        const size_t slot = (size_t)q * k;
        if (threadIdx.x == 0) s_dead = 0;
        __syncthreads();
        if (threadIdx.x < world) {
            if (!p2p_wait_flag(&mine->flags[par][threadIdx.x][q], seq, timeout_ns)) { atomicAdd(&mine->timeouts, 1u); s_dead = 1; }
        }
        __syncthreads();
        __threadfence_system();
        if (s_dead) {                        // never hand out a merge of stale items
            for (int t = threadIdx.x; t < k; t += blockDim.x) { out_ids[slot + t] = -1; out_scores[slot + t] = -INFINITY; }
            __syncthreads();
            continue;
        }
        const int total = world * k;
        int P = 32;
        while (P < total) P <<= 1;
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            GatherItem it;
            if (i < total) {
                const int r = i / k, t = i - r * k;
                const volatile GatherItem* src = &mine->items[par][r][slot + t];
                it.s = src->s;
                it.id = src->id;
                if (it.id == -1 && it.s == -INFINITY) it.id = INT64_MAX;
            } else { it.s = -INFINITY; it.id = INT64_MAX; }
            sm[i] = it;
        }
        __syncthreads();
        for (int k2 = 2; k2 <= P; k2 <<= 1) {
            for (int j = k2 >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < P; i += blockDim.x) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const bool desc = (i & k2) == 0;
                        const GatherItem a = sm[i], b = sm[ixj];
                        if (desc ? item_better(b, a) : item_better(a, b)) { sm[i] = b; sm[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        }
        for (int t = threadIdx.x; t < k; t += blockDim.x) {
            const GatherItem it = sm[t];
            const bool pad = it.id == INT64_MAX && it.s == -INFINITY;
            out_ids[slot + t] = pad ? -1 : it.id;
            out_scores[slot + t] = pad ? -INFINITY : (float)it.s;
        }
        __syncthreads();
    }
}

// Query all-gather over peer memory: rank r copied ITS contiguous slice of the query batch host -> device into its
// own region (1/world of the H2D bytes), this kernel pushes the slice to every peer, publishes one flag per CTA,
// and waits until every rank's slice has landed here.  The search that follows on the stream reads the whole batch
// from the local region.
__global__ void __launch_bounds__(256) p2p_query_allgather_kernel(P2PPeers peers, int rank, int world, unsigned int seq,
                                                                  unsigned long long timeout_ns, size_t lo4, size_t hi4 /* my slice, in float4 units */) {
    const int par = seq & 1;
    const float4* src = reinterpret_cast<const float4*>(peers.r[rank]->qbuf[par]);
    for (size_t i = lo4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = src[i];
        for (int r = 0; r < world; ++r)
            if (r != rank) reinterpret_cast<float4*>(peers.r[r]->qbuf[par])[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        volatile unsigned int* f = &peers.r[threadIdx.x]->qflags[par][rank][blockIdx.x];
        *f = seq;
    }
    P2PRegion* mine = peers.r[rank];
    // every CTA waits for a share of the (source rank, source CTA) flags; the kernel ends when all have been seen
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < world * (int)gridDim.x; i += gridDim.x * blockDim.x) {
        if (!p2p_wait_flag(&mine->qflags[par][i / gridDim.x][i % gridDim.x], seq, timeout_ns)) atomicAdd(&mine->timeouts, 1u);
    }
    __threadfence_system();
}

static void p2p_free(avs_store* s) {
    P2PState* st = (P2PState*)s->p2p_state;
    if (!st) return;
    for (int r = 0; r < P2P_MAX_WORLD; ++r)
        if (st->peers.r[r] && st->peers.r[r] != st->local) cudaIpcCloseMemHandle(st->peers.r[r]);
    if (st->local) cudaFree(st->local);
    if (st->h_timeouts) cudaFreeHost(st->h_timeouts);
    for (cudaEvent_t e : st->tev) cudaEventDestroy(e);
    cudaGetLastError();
    delete st;
    s->p2p_state = nullptr;
}

// number of bounded waits that gave up waiting for a peer (0 in a healthy job); read by avs_get_stat("p2p_timeouts")
int avs_p2p_timeouts(avs_store* s, int64_t* out) {
    P2PState* st = (P2PState*)s->p2p_state;
    *out = 0;
    if (!st || !st->local) return AVS_OK;
    unsigned int v = 0;
    AVS_CUDA(cudaMemcpy(&v, &st->local->timeouts, sizeof(v), cudaMemcpyDeviceToHost));
    *out = (int64_t)v;
    return AVS_OK;
}

// mean duration (microseconds) of the exchange kernels launched while the scan timing hook was on; -1 if none
int avs_p2p_exchange_us(avs_store* s, int64_t* out) {
    P2PState* st = (P2PState*)s->p2p_state;
    *out = -1;
    if (!st || st->tev_used == 0) return AVS_OK;
    double tot = 0.0;
    for (size_t i = 0; i < st->tev_used; ++i) {
        AVS_CUDA(cudaEventSynchronize(st->tev[2 * i + 1]));
        float ms = 0.f;
        AVS_CUDA(cudaEventElapsedTime(&ms, st->tev[2 * i], st->tev[2 * i + 1]));
        tot += ms;
    }
    *out = (int64_t)(tot / (double)st->tev_used * 1000.0 + 0.5);
    return AVS_OK;
}
void avs_p2p_timing_reset(avs_store* s) {
    P2PState* st = (P2PState*)s->p2p_state;
    if (st) st->tev_used = 0;
}

extern "C" int avs_p2p_init(avs_store* s, int rank, int world, void* handle64_out) {
    if (!s || !handle64_out || world < 1 || world > P2P_MAX_WORLD || rank < 0 || rank >= world) { avs_set_error("avs_p2p_init: bad arguments (rank %d, world %d, max world %d)", rank, world, P2P_MAX_WORLD); return AVS_E_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
    AVS_CUDA(cudaSetDevice(s->device));
    p2p_free(s);
    P2PState* st = new P2PState();
    if (cudaMalloc((void**)&st->local, sizeof(P2PRegion)) != cudaSuccess) {
        cudaGetLastError();
        delete st;
        avs_set_error("out of device memory for the peer exchange region (%zu bytes)", sizeof(P2PRegion));
        return AVS_E_NOMEM;
    }
    AVS_CUDA(cudaMemset(st->local, 0, sizeof(P2PRegion)));
    AVS_CUDA(cudaDeviceSynchronize());
    if (cudaHostAlloc((void**)&st->h_timeouts, sizeof(unsigned int), cudaHostAllocDefault) == cudaSuccess) *st->h_timeouts = 0;
    else { cudaGetLastError(); st->h_timeouts = nullptr; }
    // the exchange kernel's grid must be co-resident (its CTAs wait for peers' CTAs, never for local ones that are
    // not running yet): largest item list = 8 ranks x 256 hits -> 2048 x 16 B of dynamic shared memory
    int per_sm = 0;
    AVS_CUDA(cudaFuncSetAttribute(p2p_exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * (int)sizeof(GatherItem)));
    AVS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, p2p_exchange_merge_kernel, 256, 2048 * sizeof(GatherItem)));
    st->max_ctas = (per_sm < 1 ? 1 : per_sm) * s->num_sms;
    cudaIpcMemHandle_t h;
    AVS_CUDA(cudaIpcGetMemHandle(&h, st->local));
    memcpy(handle64_out, &h, sizeof(h));
    s->p2p_state = st;
    s->rank = rank;
    s->world = world;
    return AVS_OK;
}

extern "C" int avs_p2p_connect(avs_store* s, const void* handles, int world) {
    if (!s || !handles || !s->p2p_state || world != s->world) { avs_set_error("avs_p2p_connect: call avs_p2p_init first (and pass the same world)"); return AVS_E_STATE; }
    AVS_CUDA(cudaSetDevice(s->device));
    P2PState* st = (P2PState*)s->p2p_state;
    for (int r = 0; r < world; ++r) {
        if (r == s->rank) { st->peers.r[r] = st->local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * r, sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            avs_set_error("cudaIpcOpenMemHandle for rank %d failed: %s (no peer access between these GPUs?)", r, cudaGetErrorString(e));
            return AVS_E_CUDA;
        }
        st->peers.r[r] = (P2PRegion*)p;
    }
    st->connected = true;
    return AVS_OK;
}

extern "C" int avs_nccl_unique_id(void* out128) {
    if (!out128) { avs_set_error("avs_nccl_unique_id: NULL buffer"); return AVS_E_INVALID; }
    AVS_CHECK(nccl_load());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    AVS_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return AVS_OK;
}

extern "C" int avs_comm_init(avs_store* s, const void* unique_id128, int rank, int world) {
    if (!s || !unique_id128 || world < 1 || rank < 0 || rank >= world) { avs_set_error("avs_comm_init: bad arguments (rank %d, world %d)", rank, world); return AVS_E_INVALID; }
    AVS_CHECK(nccl_load());
    AVS_CUDA(cudaSetDevice(s->device));
    avs_comm_free(s);
    ncclUniqueId id;
    memcpy(&id, unique_id128, sizeof(id));
    ncclComm_t comm = nullptr;
    AVS_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
    s->nccl_comm = comm;
    s->rank = rank;
    s->world = world;
    return AVS_OK;
}

void avs_comm_free(avs_store* s) {
    p2p_free(s);
    if (s->nccl_comm && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)s->nccl_comm);
    s->nccl_comm = nullptr;
    s->world = 1;
    s->rank = 0;
}

// A peer that stopped answering poisoned the affected queries on the device; the pinned mirror of the counter tells the
// host without a synchronisation.  From then on the call fails loudly and the store falls back to the NCCL exchange.
static int p2p_check_health(avs_store* s, P2PState* p2p) {
    if (!p2p || !p2p->h_timeouts) return AVS_OK;
    const unsigned int now = *(volatile unsigned int*)p2p->h_timeouts;
    if (now != p2p->seen_timeouts) {
        p2p->seen_timeouts = now;
        s->opt_p2p = 0;
        avs_set_error("peer-memory exchange timed out %u time(s): a rank stopped answering; the affected queries were returned "
                      "as id -1 / -inf and this store now uses the NCCL exchange", now);
        return AVS_E_NCCL;
    }
    return AVS_OK;
}

static unsigned long long p2p_timeout_ns(const avs_store* s) {
    const long long ms = s->opt_p2p_timeout_ms > 0 ? s->opt_p2p_timeout_ms : P2P_TIMEOUT_MS_DEFAULT;
    return (unsigned long long)ms * 1000000ull;
}

static int exchange_and_merge(avs_store* s, int nq, int k, int64_t* out_ids, float* out_scores, cudaStream_t st, bool use_p2p) {
    AvsScratch& c = s->sc;
    P2PState* p2p = (P2PState*)s->p2p_state;
    if (use_p2p) {
        int P = 32;
        while (P < s->world * k) P <<= 1;
        p2p->seq += 1;
        if (p2p->seq == 0) p2p->seq = 2;   // 0 is the "never written" value of the flags; keep the parity sequence
        const int grid = nq < p2p->max_ctas ? nq : p2p->max_ctas;
        const bool timed = s->timing && p2p->tev_used < 4096;
        if (timed) {
            if (2 * p2p->tev_used >= p2p->tev.size()) {
                cudaEvent_t a, b;
                AVS_CUDA(cudaEventCreate(&a)); AVS_CUDA(cudaEventCreate(&b));
                p2p->tev.push_back(a); p2p->tev.push_back(b);
            }
            AVS_CUDA(cudaEventRecord(p2p->tev[2 * p2p->tev_used], st));
        }
        p2p_exchange_merge_kernel<<<grid, 256, (size_t)P * sizeof(GatherItem), st>>>(p2p->peers, s->rank, s->world, p2p->seq, p2p_timeout_ns(s), nq, k,
                                                                                   c.out_s64, out_ids, out_scores);
        s->st_launches++;
        AVS_CUDA(cudaGetLastError());
        if (timed) { AVS_CUDA(cudaEventRecord(p2p->tev[2 * p2p->tev_used + 1], st)); p2p->tev_used++; }
        if (p2p->h_timeouts) AVS_CUDA(cudaMemcpyAsync(p2p->h_timeouts, &p2p->local->timeouts, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        return AVS_OK;
    }
    const size_t items = (size_t)nq * k;
    const size_t cap_items = (size_t)c.nq_cap * c.k_cap;  // sized from the scratch capacities
    if (!c.gather_send || c.gather_items < cap_items || c.world_cap < s->world) {
        AVS_CUDA(cudaStreamSynchronize(st));
        cudaFree(c.gather_send); cudaFree(c.gather_recv);
        c.gather_send = c.gather_recv = nullptr;
        if (cudaMalloc(&c.gather_send, cap_items * sizeof(GatherItem)) != cudaSuccess ||
            cudaMalloc(&c.gather_recv, cap_items * sizeof(GatherItem) * s->world) != cudaSuccess) {
            cudaGetLastError();
            avs_set_error("out of device memory for the shard gather buffers");
            return AVS_E_NOMEM;
        }
        c.gather_items = cap_items;
        c.world_cap = s->world;
    }
    pack_gather_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(c.out_s64, out_ids, (GatherItem*)c.gather_send, (int64_t)items);
    s->st_launches++;
    AVS_CUDA(cudaGetLastError());
    AVS_NCCL(g_nccl.AllGather(c.gather_send, c.gather_recv, items * sizeof(GatherItem), ncclChar, (ncclComm_t)s->nccl_comm, st));
    int P = 32;
    while (P < s->world * k) P <<= 1;
    shard_merge_kernel<<<nq, 256, (size_t)P * sizeof(GatherItem), st>>>((const GatherItem*)c.gather_recv, s->world, nq, k, out_ids, out_scores);
    s->st_launches++;
    AVS_CUDA(cudaGetLastError());
    return AVS_OK;
}

extern "C" int avs_search_sharded(avs_store* s, const float* q, int nq, int k, int64_t* out_ids, float* out_scores,
                                  void* stream) {
    if (!s) { avs_set_error("avs_search_sharded: NULL store"); return AVS_E_INVALID; }
    P2PState* p2p = (P2PState*)s->p2p_state;
    AVS_CHECK(p2p_check_health(s, p2p));
    const bool use_p2p = p2p && p2p->connected && s->opt_p2p && (size_t)nq * k <= P2P_ITEMS_PER_SRC && nq <= P2P_MAX_NQ;
    if (!use_p2p && !s->nccl_comm) { avs_set_error("avs_search_sharded: neither avs_comm_init nor avs_p2p_connect has been called on this store"); return AVS_E_STATE; }
    if (s->filter) { avs_set_error("avs_search_sharded: row filters are not supported on a sharded store"); return AVS_E_STATE; }
    if (k > AVS_MAX_KPRIME) { avs_set_error("avs_search_sharded: limit %d above %d is served on unsharded stores only", k, AVS_MAX_KPRIME); return AVS_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    // local exact top-k; ids/scores land in the caller's buffers first and are then replaced
    AVS_CHECK(avs_search_local(s, q, nq, k, out_ids, out_scores, nullptr, st));
    if (nq == 0) return AVS_OK;
    return exchange_and_merge(s, nq, k, out_ids, out_scores, st, use_p2p);
}

// End to end from HOST buffers on a sharded store.  Every rank holds the same host query batch; with the peer regions
// connected each rank copies only ITS 1/world slice host -> device and the slices are all-gathered over NVLink by
// p2p_query_allgather_kernel (one PCIe copy of the batch per node instead of `world`); otherwise every rank copies the
// whole batch.  Local search, exchange + merge, one D2H copy of the merged hits, synchronised on return.
extern "C" int avs_search_sharded_host(avs_store* s, const float* q_host, int nq, int k, int64_t* out_ids_host,
                                       float* out_scores_host) {
    if (!s) { avs_set_error("avs_search_sharded_host: NULL store"); return AVS_E_INVALID; }
    if (nq < 0 || (nq > 0 && (!q_host || !out_ids_host || !out_scores_host))) { avs_set_error("avs_search_sharded_host: NULL buffer"); return AVS_E_INVALID; }
    if (nq == 0) return AVS_OK;
    P2PState* p2p = (P2PState*)s->p2p_state;
    AVS_CHECK(p2p_check_health(s, p2p));
    const bool use_p2p = p2p && p2p->connected && s->opt_p2p && (size_t)nq * k <= P2P_ITEMS_PER_SRC && nq <= P2P_MAX_NQ;
    if (!use_p2p && !s->nccl_comm) { avs_set_error("avs_search_sharded_host: neither avs_comm_init nor avs_p2p_connect has been called on this store"); return AVS_E_STATE; }
    if (s->filter) { avs_set_error("avs_search_sharded_host: row filters are not supported on a sharded store"); return AVS_E_STATE; }
    if (k > AVS_MAX_KPRIME) { avs_set_error("avs_search_sharded_host: limit %d above %d is served on unsharded stores only", k, AVS_MAX_KPRIME); return AVS_E_INVALID; }
    AVS_CUDA(cudaSetDevice(s->device));
    AVS_CHECK(avs_host_staging_reserve(s, nq, k));
    AvsScratch& c = s->sc;
    cudaStream_t st = 0;
    const size_t items = (size_t)nq * k;
    int64_t* d_ids = c.d_ids;
    float* d_scores = reinterpret_cast<float*>(c.d_ids + 2 * items);
    const size_t total_f = (size_t)nq * s->dim;
    const float* q_dev = c.h2d_q;
    // queries leave from the caller's buffer when it is pinned, else through the pinned stage (only the bytes this rank copies)
    cudaPointerAttributes pa;
    const bool pinned = cudaPointerGetAttributes(&pa, q_host) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const bool gather_q = use_p2p && s->world > 1 && (s->dim % 4 == 0) && total_f <= P2P_Q_FLOATS && nq >= s->world;
    if (gather_q) {
        p2p->qseq += 1;
        if (p2p->qseq == 0) p2p->qseq = 2;
        const int par = p2p->qseq & 1;
        const int per = (nq + s->world - 1) / s->world;
        const int lo = s->rank * per < nq ? s->rank * per : nq, hi = lo + per < nq ? lo + per : nq;
        float* qbuf = p2p->local->qbuf[par];
        if (hi > lo) {
            const size_t off = (size_t)lo * s->dim, nb = (size_t)(hi - lo) * s->dim * sizeof(float);
            const float* src = q_host + off;
            if (!pinned) { memcpy(c.h_q + off, q_host + off, nb); src = c.h_q + off; }
            AVS_CUDA(cudaMemcpyAsync(qbuf + off, src, nb, cudaMemcpyHostToDevice, st));
        }
        p2p_query_allgather_kernel<<<P2P_Q_CTAS, 256, 0, st>>>(p2p->peers, s->rank, s->world, p2p->qseq, p2p_timeout_ns(s),
                                                               (size_t)lo * s->dim / 4, (size_t)hi * s->dim / 4);
        s->st_launches++;
        AVS_CUDA(cudaGetLastError());
        q_dev = qbuf;
    } else {
        const float* src = q_host;
        if (!pinned) { memcpy(c.h_q, q_host, total_f * sizeof(float)); src = c.h_q; }
        AVS_CUDA(cudaMemcpyAsync(c.h2d_q, src, total_f * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    AVS_CHECK(avs_search_local(s, q_dev, nq, k, d_ids, d_scores, nullptr, st));
    AVS_CHECK(exchange_and_merge(s, nq, k, d_ids, d_scores, st, use_p2p));
    AVS_CUDA(cudaMemcpyAsync(c.h_out, d_ids, items * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    AVS_CUDA(cudaMemcpyAsync(c.h_out + 2 * items, d_scores, items * sizeof(float), cudaMemcpyDeviceToHost, st));
    AVS_CUDA(cudaStreamSynchronize(st));
    AVS_CHECK(p2p_check_health(s, p2p));
    memcpy(out_ids_host, c.h_out, items * sizeof(int64_t));
    memcpy(out_scores_host, c.h_out + 2 * items, items * sizeof(float));
    return AVS_OK;
}
