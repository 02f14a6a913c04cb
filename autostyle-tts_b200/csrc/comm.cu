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
// Every rank exposes one exchange region through CUDA IPC.  One kernel per search, one CTA per query:
//   1. store this rank's k (score, id) items of the query straight into slot [rank] of EVERY peer's region
//      (remote stores over NVLink) and of its own,
//   2. __threadfence_system, then publish flag[rank][q] = seq in every region,
//   3. spin until the local region shows seq for the query from every rank,
//   4. merge world*k items by (score desc, id asc) and write the global top-k.
// A CTA pushes before it waits and never depends on another local CTA, so any scheduling order completes.
// Two parities alternate between searches: a rank can only be one search ahead of its slowest peer (it needs
// that peer's push to finish its own merge), so a region is never overwritten while it is still being read.
// ---------------------------------------------------------------------------------------------
#define P2P_MAX_WORLD 8
#define P2P_ITEMS_PER_SRC (1 << 18)          // (score, id) items per source rank per parity (4 MiB)
#define P2P_MAX_NQ (1 << 14)

struct P2PRegion {                            // layout of one rank's exchange region
    GatherItem items[2][P2P_MAX_WORLD][P2P_ITEMS_PER_SRC];
    unsigned int flags[2][P2P_MAX_WORLD][P2P_MAX_NQ];
    unsigned int timeouts;
};
struct P2PPeers { P2PRegion* r[P2P_MAX_WORLD]; };

struct P2PState {
    P2PRegion* local = nullptr;
    P2PPeers peers{};
    bool connected = false;
    unsigned int seq = 0;
};

__global__ void __launch_bounds__(256) p2p_exchange_merge_kernel(P2PPeers peers, int rank, int world, unsigned int seq,
                                                                 int nq, int k, const double* __restrict__ s64,
                                                                 int64_t* __restrict__ out_ids, float* __restrict__ out_scores) {
    extern __shared__ unsigned char raw[];
    GatherItem* sm = reinterpret_cast<GatherItem*>(raw);
    const int q = blockIdx.x, par = seq & 1;
    const size_t slot = (size_t)q * k;
    // 1. push my items for this query into every region (mine included)
    for (int i = threadIdx.x; i < world * k; i += blockDim.x) {
        const int r = i / k, t = i - r * k;
        GatherItem it;
        it.s = s64[slot + t];
        it.id = out_ids[slot + t];
        peers.r[r]->items[par][rank][slot + t] = it;
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish
    if (threadIdx.x < world) {
        volatile unsigned int* f = &peers.r[threadIdx.x]->flags[par][rank][q];
        *f = seq;
    }
    // 3. wait for every rank's items of this query in MY region (bounded spin: a missing peer must not hang the GPU)
    P2PRegion* mine = peers.r[rank];
    if (threadIdx.x < world) {
        volatile unsigned int* f = &mine->flags[par][threadIdx.x][q];
        long long spins = 0;
        while (*f != seq) {
            if (++spins > (1ll << 28)) { atomicAdd(&mine->timeouts, 1u); break; }
        }
    }
    __syncthreads();
    __threadfence_system();
    // 4. merge
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
}

static void p2p_free(avs_store* s) {
    P2PState* st = (P2PState*)s->p2p_state;
    if (!st) return;
    for (int r = 0; r < P2P_MAX_WORLD; ++r)
        if (st->peers.r[r] && st->peers.r[r] != st->local) cudaIpcCloseMemHandle(st->peers.r[r]);
    if (st->local) cudaFree(st->local);
    cudaGetLastError();
    delete st;
    s->p2p_state = nullptr;
}

// number of bounded spins that gave up waiting for a peer (0 in a healthy job); read by avs_get_stat("p2p_timeouts")
int avs_p2p_timeouts(avs_store* s, int64_t* out) {
    P2PState* st = (P2PState*)s->p2p_state;
    *out = 0;
    if (!st || !st->local) return AVS_OK;
    unsigned int v = 0;
    AVS_CUDA(cudaMemcpy(&v, &st->local->timeouts, sizeof(v), cudaMemcpyDeviceToHost));
    *out = (int64_t)v;
    return AVS_OK;
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

extern "C" int avs_search_sharded(avs_store* s, const float* q, int nq, int k, int64_t* out_ids, float* out_scores,
                                  void* stream) {
    if (!s) { avs_set_error("avs_search_sharded: NULL store"); return AVS_E_INVALID; }
    P2PState* p2p = (P2PState*)s->p2p_state;
    const bool use_p2p = p2p && p2p->connected && s->opt_p2p && (size_t)nq * k <= P2P_ITEMS_PER_SRC && nq <= P2P_MAX_NQ;
    if (!use_p2p && !s->nccl_comm) { avs_set_error("avs_search_sharded: neither avs_comm_init nor avs_p2p_connect has been called on this store"); return AVS_E_STATE; }
    if (s->filter) { avs_set_error("avs_search_sharded: row filters are not supported on a sharded store"); return AVS_E_STATE; }
    cudaStream_t st = (cudaStream_t)stream;
    // local exact top-k; ids/scores land in the caller's buffers first and are then replaced
    AVS_CHECK(avs_search_local(s, q, nq, k, out_ids, out_scores, nullptr, st));
    if (nq == 0) return AVS_OK;
    AvsScratch& c = s->sc;
    if (use_p2p) {
        int P = 32;
        while (P < s->world * k) P <<= 1;
        p2p->seq += 1;
        if (p2p->seq == 0) p2p->seq = 2;   // 0 is the "never written" value of the flags; keep the parity sequence
        p2p_exchange_merge_kernel<<<nq, 256, (size_t)P * sizeof(GatherItem), st>>>(p2p->peers, s->rank, s->world, p2p->seq, nq, k,
                                                                                 c.out_s64, out_ids, out_scores);
        s->st_launches++;
        AVS_CUDA(cudaGetLastError());
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
