/* avs.h — C-ABI of the B200-native exact (FLAT) vector store.
 *
 * This is the drop-in boundary for the one hot path of AutoStyle-TTS: the
 * brute-force cosine / inner-product top-k search the reference performs through
 * `pymilvus.MilvusClient` on Milvus Lite.  Every entry point below names the
 * reference interface it replaces (file:line under /root/reference).  Plain C:
 * pointers and sizes only, no C++ exceptions, no Python or torch types.
 *
 * Conventions
 *   - every function returns AVS_OK (0) or a negative AVS_E_* code; the message
 *     for the calling thread's last failure is avs_last_error().
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - one store <-> one device.  Not safe for concurrent insert; concurrent
 *     searches need distinct stores or external serialisation (scratch buffers
 *     are per store).
 *   - scores are SIMILARITIES (larger = better) exactly like the reference's
 *     `hit['distance']` (/root/reference/milvus/search_embeddings.py:54,
 *     /root/reference/output_emb/search_results.json).
 *   - result order: (score desc, id asc), evaluated in float64 from the fp32
 *     master rows; slots past min(k, count) hold id -1 / score -inf.
 */
#ifndef AVS_H_
#define AVS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct avs_store avs_store;

enum { AVS_METRIC_COSINE = 0, AVS_METRIC_IP = 1 };

enum {
    AVS_OK = 0,
    AVS_E_INVALID = -1,   /* bad argument */
    AVS_E_CUDA = -2,      /* CUDA runtime / driver failure */
    AVS_E_NOMEM = -3,     /* capacity exceeded or allocation failed */
    AVS_E_NCCL = -4,      /* NCCL missing or failed */
    AVS_E_STATE = -5      /* call not valid in this state (e.g. sharded search before comm_init) */
};

/* Replaces `MilvusClient.create_collection(collection_name=, dimension=)`
 * (/root/reference/milvus/RAG.py:54-57) and the schema variant
 * (/root/reference/milvus/insert_embeddings.py:52-63): allocates the device-resident
 * store — fp32 master rows, the bf16 scan copy (L2-normalised for COSINE), inverse
 * norms and primary keys — for up to `capacity` rows of `dim` floats on `device`. */
int avs_create(int device, int dim, int metric, int64_t capacity, avs_store** out);

/* Replaces `MilvusClient.drop_collection` (/root/reference/milvus/RAG.py:50). */
int avs_destroy(avs_store* s);

/* Grows the store to at least `capacity` rows (device-to-device copy). */
int avs_reserve(avs_store* s, int64_t capacity);

/* Replaces `MilvusClient.insert(collection_name=, data=[{"id","vector",...}])`
 * (/root/reference/milvus/RAG.py:541-544, /root/reference/milvus/insert_embeddings.py:519).
 * `rows` = n x dim fp32, host or device memory (detected); `ids` = n int64 primary
 * keys, host or device, or NULL for auto ids (row index, the `auto_id=True` case of
 * /root/reference/milvus/insert_embeddings.py:53).  Rows are stored un-normalised in
 * the master exactly as given; the normalise-on-insert kernel builds the bf16 copy. */
int avs_insert(avs_store* s, const float* rows, const int64_t* ids, int64_t n, void* stream);

/* Benchmark/test utility (no reference counterpart): appends n synthetic unit-norm
 * rows generated ON DEVICE by the counter-based generator documented in DESIGN.md;
 * row r of the stream `seed` is a pure function of (seed, first_row + r), replayable
 * on the host bit for bit.  ids = id_base + first_row + r. */
int avs_fill_synthetic(avs_store* s, uint64_t seed, int64_t first_row, int64_t n,
                       int64_t id_base, void* stream);

/* Number of rows stored (the reference reads it through `client.get_collection_stats`). */
int64_t avs_count(const avs_store* s);
int avs_dim(const avs_store* s);
int avs_metric(const avs_store* s);

/* Copies master rows [first, first+n) (fp32, as inserted) to host or device memory. */
int avs_get_rows(avs_store* s, int64_t first, int64_t n, float* out, void* stream);
int avs_get_ids(avs_store* s, int64_t first, int64_t n, int64_t* out, void* stream);

/* Bulk snapshot / restore of a device store (the reference's persistence is the Milvus Lite
 * file re-opened by every script, /root/reference/milvus/search.py:197-210, one blob per row - kept for small collections
 * by the Python layer).  File = 64-byte header | ids | fp32 master rows as inserted; the bf16 scan copy and the norms are
 * rebuilt by the normalise-on-insert kernel on load.  Synchronous. */
int avs_save(avs_store* s, const char* path);
int avs_load(const char* path, int device, avs_store** out);

/* Replaces `MilvusClient.search(collection_name, data=[vec,...], limit=k, ...)`
 * (/root/reference/milvus/search_embeddings.py:15-22, /root/reference/milvus/RAG.py:383-390,
 * /root/reference/milvus/search_json.py:247-254, /root/reference/src/search_milvus.py:139-146).
 * `q` = nq x dim fp32 on the store's device (un-normalised queries are fine: COSINE
 * normalises inside, /root/reference/milvus/RAG.py:264 sends norm~40 vectors);
 * out_ids [nq,k] int64 / out_scores [nq,k] fp32 device buffers owned by the caller.
 * out_rows (nullable) receives the local row index of every hit (metadata lookup key).
 * Asynchronous on `stream`. 1 <= k <= 16384 (MilvusClient's own ceiling).  Limits up to 256 run the fused
 * bf16 scan + float64 rescoring pipeline; larger ones (the reference never asks for more than 5,
 * /root/reference/milvus/search_embeddings.py:64) are served exactly, and more slowly, straight from the fp32
 * master: three passes over it per 8 queries.  Sharded stores (avs_search_sharded*) take k <= 256. */
int avs_search(avs_store* s, const float* q, int nq, int k,
               int64_t* out_ids, float* out_scores, int64_t* out_rows, void* stream);

/* Replaces the `filter=` argument of `MilvusClient.search` (/root/reference/milvus/RAG.py:387; always None in the
 * reference): `bitmap_host` holds one bit per stored row (bit r of word r/32), set = the row may be returned.  The
 * host evaluates the scalar expression on its metadata; the scan kernels test the bit on their (rare) accept
 * path, so thresholds, certificate and repair all see the allowed rows only.  NULL clears the filter. */
int avs_set_filter(avs_store* s, const uint32_t* bitmap_host, int64_t n_bits);

/* Same search end to end from HOST buffers: H2D of the queries, the device pipeline,
 * D2H of ids/scores/rows, synchronised on return.  This is the call the Python
 * MilvusClient drop-in makes for list / ndarray queries. */
int avs_search_host(avs_store* s, const float* q_host, int nq, int k,
                    int64_t* out_ids_host, float* out_scores_host, int64_t* out_rows_host);

/* Multi-GPU (row-sharded store, one process per GPU).
 * avs_nccl_unique_id fills 128 bytes on rank 0; the caller broadcasts them (the
 * Python layer uses torch.distributed) and every rank calls avs_comm_init. */
int avs_nccl_unique_id(void* out128);
int avs_comm_init(avs_store* s, const void* unique_id128, int rank, int world);
/* Local exact top-k on this rank's shard -> one ncclAllGather of [nq,k] (score f64,
 * id i64) -> merge by (score desc, id asc); every rank receives the global result. */
int avs_search_sharded(avs_store* s, const float* q, int nq, int k,
                       int64_t* out_ids, float* out_scores, void* stream);
/* Same sharded search end to end from HOST buffers (every rank passes the same batch): with the peer regions connected
 * each rank copies only its 1/world slice of the queries host -> device and the slices are all-gathered over NVLink
 * peer memory; local search, exchange + merge, D2H of the merged hits; synchronised on return.  This is what the
 * reference's per-utterance `client.search(data=[emb])` loop (/root/reference/milvus/search_json.py:382-411) becomes on a
 * row-sharded store. */
int avs_search_sharded_host(avs_store* s, const float* q_host, int nq, int k,
                            int64_t* out_ids_host, float* out_scores_host);
/* Optional fused exchange over NVLink peer memory (replaces the ncclAllGather + merge of
 * avs_search_sharded by ONE kernel that stores each rank's top-k straight into its peers' memory, flags
 * it, waits for the peers' items and merges).  avs_p2p_init allocates this rank's exchange region and
 * returns its 64-byte cudaIpcMemHandle; the caller all-gathers the handles (world x 64 bytes, rank order)
 * and passes them to avs_p2p_connect.  world <= 8, nq*k <= 262144 per call (larger calls use NCCL). */
int avs_p2p_init(avs_store* s, int rank, int world, void* handle64_out);
int avs_p2p_connect(avs_store* s, const void* handles, int world);

/* Options: "scan_path" 0=auto 1=gemv 2=gemm; "hybrid" 0|1 (auto mode, <= 8 queries: dense warp-dot level, tensor-core
 * scan for the later levels; default 1 - only used where the boot level below does not apply); "boot" 0|1 (tensor-core
 * scan: the threshold-free level keeps the 8 best keys of every half row group instead of storing every key, so it
 * may visit up to 32 K rows and every level of every batch size runs in the ONE persistent launch; default 1);
 * "p2p_timeout_ms" (wall-clock bound of a wait for a peer in the exchange kernels, default 120000);
 * "oversample" K' override (0=auto); "gemm_min_batch";
 * "levels_ratio"; "cta_group" 1|2 (tensor-core scan variant) and "cta_group_small" 1|2 (the variant for batches of at most
 * 128 queries; default 1: M = 128 queries per CTA, half the padded MMA work); schedule knobs "fine_ratio",
 * "fine_min_batch", "final_sigma", "coarse_sigma"; "p2p_merge" 0|1; "pdl" 0|1 (programmatic dependent launch between
 * the kernels of a search, default 1); "trace" 0|1; "force_repair" (testing: 1 = force the wide-rescoring stage, 2 = force
 * the exact scan as well); "eps_rule" 0|1 (testing: the last threshold is kept 2.5 eps under the k-th scan score - the
 * engine switches this on by itself once a store has needed an exact repair).  Measurement knobs kept for A/B runs:
 * "boot2_ratio" (boot level directly in front of the final one up to this stride ratio; 0 = never, the default: measured
 * negative), "finalize_threads" 0|256|512|1024, "gemm_dense_rows" / "dense_rows" (row caps of the threshold-free level of the
 * tensor-core / warp-dot path when no boot level is built). */
int avs_set_option(avs_store* s, const char* key, int64_t value);
/* Counters since creation: "kernel_launches", "searches", "queries", "wide_rescored_queries" (certificate
 * reached after rescoring the whole collected set), "repaired_queries" (exact float64 scan needed),
 * "uncertified_queries", "p2p_timeouts", "exchange_us" (mean duration of the peer-memory exchange kernel while
 * avs_scan_timing is on), "last_kprime", "last_levels", "last_boot", "last_scan_path", "last_final_rows";
 * "barrier_timeouts" (a grid barrier of a persistent kernel gave up: must stay 0);
 * "last_uncertified": queries of the last avs_search_host call whose top-k could not be proven exact - read without a
 * device synchronisation.  Every query whose certificate fails is settled by the exact float64 repair (all of them, in
 * groups; a query without a usable lower bound gets one from a histogram search).  A query is left uncertified only when
 * more rows than its repair slice holds (4096, or pool / flagged-queries if thousands of queries fail at once, never
 * fewer than 256) have scores within 2^-15 relative of its k-th best - masses of (near-)duplicate rows.  Its hits are then
 * the best found by the earlier stages: possibly not the exact top-k, and reported, never silent. */
int avs_get_stat(avs_store* s, const char* key, int64_t* out);

/* Timing hook for bench.py: when enabled, CUDA events bracket the dominant scan
 * kernel of each search on its own stream; returns the mean duration in ms of the
 * launches since the last reset and their count. */
int avs_scan_timing(avs_store* s, int enable_reset, double* mean_ms, int64_t* launches);

const char* avs_last_error(void);
const char* avs_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AVS_H_ */
