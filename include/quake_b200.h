/*
 * quake_b200.h -- C ABI of the B200 (sm_100a) partitioned-IVF search hot path.
 *
 * This is the drop-in boundary for the Quake hot path. The reference has no
 * FFI for this path: its seam is the pybind11 module `quake._bindings`
 * (/root/reference/src/cpp/bindings/wrap.cpp:48), below which the functions listed here are plain C++
 * calls. Each entry point cites the reference function it replaces. Signatures carry only plain
 * pointers and sizes (device pointers unless stated otherwise) and a `cudaStream_t` passed as
 * `void*`; no torch types. All calls are asynchronous on `stream` unless stated otherwise; they
 * return 0 on success and a non-zero code on failure, in which case `qk_last_error()` describes
 * the failure (same role as the reference's std::runtime_error / std::invalid_argument messages,
 * src/cpp/include/common.h:154).
 *
 * There is NO CPU fallback behind any of these functions.
 */
#ifndef QUAKE_B200_H
#define QUAKE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same numeric values as faiss::MetricType, which the reference writes into metadata.txt
 * (src/cpp/src/quake_index.cpp:187). */
#define QK_METRIC_INNER_PRODUCT 0
#define QK_METRIC_L2 1

#define QK_OK 0
#define QK_ERR_INVALID_ARGUMENT 1
#define QK_ERR_CUDA 2
#define QK_ERR_WORKSPACE 3
#define QK_ERR_UNSUPPORTED 4

/* Largest k (neighbours per query) the scan path accepts. */
#define QK_MAX_K 2048

/*
 * Device-resident partition store: the B200 counterpart of faiss::DynamicInvertedLists +
 * IndexPartition (src/cpp/include/dynamic_inverted_list.h:33, src/cpp/include/index_partition.h:24-29).
 * Every partition ("list") is a contiguous run of rows in one arena; a list is cut into scan
 * segments of at most `QK_SEGMENT_ROWS` rows so that oversized lists (a flat index is one list)
 * still spread over the whole GPU.
 */
typedef struct qk_store {
    const float*   vectors;      /* [rows x pitch] float32, row-major, 16-byte aligned rows          */
    const int64_t* ids;          /* [rows] int64 vector ids; NULL => the arena row index is the id  */
    int64_t        pitch;        /* floats per row, >= d, multiple of 4; padding must be zero       */
    int32_t        d;            /* vector dimension                                                 */
    int32_t        num_lists;    /* number of list slots                                             */
    const int32_t* list_seg0;    /* [num_lists] first segment of each list                           */
    const int32_t* list_nseg;    /* [num_lists] number of segments of each list (0 for an empty list)*/
    int32_t        num_segments; /* total number of segments                                         */
    int32_t        max_list_segments; /* max over lists of list_nseg (host-known)                    */
    const int64_t* seg_row0;     /* [num_segments] first arena row of the segment                    */
    const int32_t* seg_rows;     /* [num_segments] rows in the segment (1..QK_SEGMENT_ROWS)          */
    float          max_row_norm; /* upper bound on the L2 norm of any stored row (see qk_max_row_norm)*/
    const float*   row_norms;    /* [rows] squared L2 norm of every row (see qk_row_sqnorms); l2 only */
    int64_t        num_rows;     /* rows allocated behind `vectors` (bounds of the TMA tensor map)       */
    int64_t        flat_row0;    /* single-list stores (num_lists == 1): first arena row of the list ...  */
    int64_t        flat_rows;    /* ... and its length (host-known copy; 0 otherwise)                     */
    int32_t        max_segment_rows; /* largest seg_rows entry (host-known; 0 = unknown)                   */
    int32_t        filter_terms; /* tensor-core filter precision for this store: 0 or 3 = 3xTF32 (error ~2^-20, bounded
                                   * 2^-17), 2 = 2xTF32 (a_hi (b_hi + b_lo), bound 2^-10): faster, and still exact results
                                   * -- the refine step's proof sends a query to the exact re-scan whenever the looser
                                   * filter cannot separate its k-th neighbour from the rest                             */
} qk_store_t;

#define QK_SEGMENT_ROWS 4096

/* ---- library info ------------------------------------------------------------------------- */
const char* qk_version(void);
const char* qk_last_error(void);
/* Kernels launched by this library since it was loaded (host-side launches; a CUDA-graph replay launches the kernels
 * that were counted while it was captured). Measurement aid: bench.py's gpu_launches. */
long long qk_launch_count(void);
/* Host call. Fails unless the current device is compute capability 10.x. */
int qk_device_check(int* sm_count, int* cc_major, int* cc_minor);

/* ---- partition scan + top-k ----------------------------------------------------------------
 * Replaces QueryCoordinator::scan_partitions -> serial_scan / batched_serial_scan
 * (src/cpp/src/query_coordinator.cpp:659-673, 471-611, 675-799), i.e. scan_list /
 * batched_scan_list + TopkBuffer (src/cpp/include/list_scanning.h:241-366, 41-204) over the probed
 * lists of every query, with the reference's output conventions: ids int64 / distances float32
 * [Q x k], best first, L2 distances are Euclidean (sqrt) (list_scanning.h:260), missing slots are
 * id -1 and +inf (l2) / -inf (ip) (query_coordinator.cpp:589-601).
 *
 * probe_lists: [Q x nprobe] int32 list slots in probe order, -1 = skip (query_coordinator.cpp:540). May be NULL for a
 *              store with ONE list and nprobe = 1 ("flat mode": a flat index, the centroid list of the coarse scan --
 *              every query scans the whole list, query_coordinator.cpp:624-626); the work items are then generated
 *              inside the scan kernel and no grouping kernels run.
 * out_rows   : optional [Q x k] int64 arena row of each result (-1 where padded).
 * stats      : optional device int32[8]: {queries that took the exact re-scan path, largest number of
 *              candidates any query appended, total candidates appended (low, high 32 bits), queries whose proof
 *              would have failed under the 2xTF32 filter's error bound (3xTF32 scans of multi-list stores only;
 *              see qk_store_t.filter_terms), 3 reserved}.
 */
size_t qk_scan_workspace_bytes(const qk_store_t* store, int64_t num_queries, int nprobe, int k);
int qk_scan_partitions(const qk_store_t* store,
                       const float* queries, int64_t num_queries, int64_t query_pitch,
                       const int32_t* probe_lists, int nprobe,
                       int metric, int k,
                       int64_t* out_ids, float* out_distances, int64_t* out_rows,
                       void* workspace, size_t workspace_bytes,
                       int32_t* stats, void* stream);

/* Two-level fixed-nprobe search in one call. Replaces QueryCoordinator::search for recall_target <= 0
 * (src/cpp/src/query_coordinator.cpp:612-657): coarse scan of the flat parent (top-min(nprobe, nlist) centroids per
 * query, :628-645) -> partition ids -> list slots (id_to_slot, the partitions_.at(pid) lookup) -> scan of the probed
 * lists -> top-k. `parent` must be a single-list store whose ids are the partition ids. shard_world > 1 scans only the
 * partitions with id % shard_world == shard_rank (lists sharded over GPUs, partition_manager.cpp:599-602): the result
 * is then this shard's PARTIAL top-k. out_probe_ids: optional [Q x min(nprobe, nlist)] int64, the probed partition
 * ids as a SET -- exactly the reference's nprobe nearest centroids (membership is settled in exact arithmetic wherever
 * the filter's error bound leaves a doubt), the nearest by filter score first, the rest in no guaranteed order: the
 * partition scan does not depend on it and the hit window of the maintenance policy counts per partition. */
size_t qk_search_ivf_workspace_bytes(const qk_store_t* parent, const qk_store_t* store, int64_t num_queries,
                                     int nprobe, int k);
int qk_search_ivf(const qk_store_t* parent, const qk_store_t* store,
                  const int32_t* id_to_slot, int64_t table_size,
                  const float* queries, int64_t num_queries, int64_t query_pitch,
                  int nprobe, int metric, int k, int shard_rank, int shard_world,
                  int64_t* out_ids, float* out_distances, int64_t* out_probe_ids,
                  void* workspace, size_t workspace_bytes, int32_t* stats, void* stream);

/* Measurement aid (bench.py): while profiling is on, every qk_scan_partitions call records a CUDA-event
 * pair on its stream right around the filter ("scan") kernel -- the dominant kernel of the path. Read the
 * records after synchronising. Not thread-safe; off by default. */
int qk_profile_begin(int max_records);
int qk_profile_count(void);
int qk_profile_read(int index, float* ms, int64_t* queries, int* nprobe, int* k);
int qk_profile_end(void);

/* Map the ids returned by the coarse (parent) scan -- partition ids -- to list slots of the child
 * store. id_to_slot is a dense device table of `table_size` entries; ids outside it or mapped to
 * a negative value become -1. Replaces the unordered_map lookup partitions_.at(pid)
 * (src/cpp/src/dynamic_inverted_list.cpp) on the search path. */
int qk_map_ids_to_slots(const int64_t* ids, int64_t n, const int32_t* id_to_slot, int64_t table_size,
                        int32_t* out_slots, void* stream);

/* Upper bound on the row norms of [n x pitch] rows, written to a device float (max-reduced into
 * *out, which the caller initialises, e.g. to 0). Used for the filter/refine safety bound. */
int qk_max_row_norm(const float* rows, int64_t n, int64_t pitch, int d, float* out, void* stream);

/* out[i] = ||rows[i]||^2 (fp32): the per-row term of the l2 filter score ||v||^2 - 2<q,v>. The store keeps it
 * beside the vectors (4 B per row) so that the scan kernel streams each vector exactly once. */
int qk_row_sqnorms(const float* rows, int64_t n, int64_t pitch, int d, float* out, void* stream);

/* Merge S partial results per query (multi-GPU shards, or APS rounds) into one top-k.
 * Replaces the global TopkBuffer::batch_add merge of per-core partial results
 * (src/cpp/src/query_coordinator.cpp:167-173, 752-758). parts are [S x Q x k]; padded entries have
 * id -1. Order: by distance (ascending l2 / descending ip), ties by ascending id. */
int qk_merge_topk(const float* part_distances, const int64_t* part_ids, int num_parts,
                  int64_t num_queries, int k, int metric,
                  int64_t* out_ids, float* out_distances, void* stream);

/* ---- shard exchange over NVLink peer memory (lists sharded over the GPUs of one box) ------------------
 * Replaces, across GPUs, the merge of per-core partial results into the global TopkBuffer
 * (src/cpp/src/query_coordinator.cpp:167-173): every rank pushes its [Q x k] partial top-k straight into every
 * peer's buffer (remote stores + one flag per CTA), waits for the peers' pushes and merges -- ONE kernel per rank, no
 * NCCL call on the query path; the result is identical on every rank and bit-identical to the unsharded search.
 * A peer buffer is cudaMalloc'ed by qk_peer_alloc (HOST call; also returns its 64-byte CUDA IPC handle, which the
 * caller distributes to the other ranks, e.g. with one all-gather at set-up time) and opened on the other ranks with
 * qk_peer_open. peer_buffers: HOST array of `world` device pointers, entry `rank` being the local buffer; all of
 * qk_peer_buffer_bytes(Q, k, world) bytes. Collective: every rank must call it the same number of times with the
 * same Q, k, world. The exchange counter lives in the buffer, so the call can be captured in a CUDA graph. */
size_t qk_peer_buffer_bytes(int64_t num_queries, int k, int world);
int qk_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_64);
int qk_peer_open(const void* ipc_handle_64, void** dev_ptr);
int qk_peer_close(void* dev_ptr);
int qk_peer_free(void* dev_ptr);
int qk_exchange_merge_topk(const int64_t* ids, const float* distances, int64_t num_queries, int k, int metric,
                           int rank, int world, void* const* peer_buffers,
                           int64_t* out_ids, float* out_distances, void* stream);

/* ---- Adaptive Partition Scanning (recall_target > 0) ----------------------------------------------
 * Replaces the APS part of QueryCoordinator::serial_scan (src/cpp/src/query_coordinator.cpp:521-579) and the
 * geometry it calls (src/cpp/include/geometry.h).
 *
 * qk_host_beta_table: HOST call; table[1001] = I_x((d+1)/2, 1/2) at x = i/1000 (geometry.h:163-186; the caller
 * uploads it once per dimension).
 * qk_aps_boundary_distances: compute_boundary_distances (geometry.h:57-113) for every query against its m
 * rank-ordered candidate centroids. cand_rows [Q x m]: arena rows of the candidates in `centroids` (the parent
 * store's vectors), -1 where the coarse scan returned fewer; out [Q x m], entry 0 (and invalid ones) = -1.
 * qk_aps_advance: one ROUND of the reference's per-query loop over probe ranks p0 .. p0+R-1 for the queries
 * listed in `active`: round_ids / round_distances [num_active x R x k] hold the top-k of every (query, rank)
 * partition scan of the round (one qk_scan_partitions call with one probe per pseudo-query). Per query: merge,
 * k-th distance, recall profile (compute_recall_profile, geometry.h:345-407) when the radius moved by more than
 * recompute_threshold, stop when sum(probs[0..p-1]) >= recall_target. State arrays are indexed by query:
 * run_ids/run_distances [Q x k] (padded -1 / +-inf), run_count, radius (initialise to +-1e6,
 * query_coordinator.cpp:524-527), have_probs, probs [Q x m], done, scanned (partitions scanned so far).
 * still_active: device int32, number of listed queries that need another round. k <= 1024. */
/* qk_aps_thresholds: per listed query, the filter-key threshold under which every row that can still enter its
 * running top-k must fall (k-th distance so far, widened by the filter's error bound; +inf while fewer than k
 * results are held). qk_scan_collect: the partition scan with those thresholds given and FIXED -- every row under
 * them is kept, refined in the reference's exact arithmetic and handed out grouped by the probe rank of its list,
 * each group best first and cut at k: out_ids / out_distances [Q x nprobe x k], out_cnt [Q x nprobe] entries per
 * group, out_overflow [Q] = 1 when a query's candidate buffer overflowed (its groups are then unusable). One call
 * replaces R pseudo-queries with a full top-k each (the per-list results serial_scan merges one by one,
 * query_coordinator.cpp:537-580). Lists must be single-segment (<= QK_SEGMENT_ROWS rows); workspace as for
 * qk_scan_partitions. */
int qk_aps_thresholds(const int32_t* active, int64_t num_active, const float* queries, int64_t query_pitch, int d,
                      const float* run_distances, const int32_t* run_count, int k, int metric,
                      float max_row_norm, int filter_terms, uint32_t* out_keys, void* stream);
int qk_scan_collect(const qk_store_t* store, const float* queries, int64_t num_queries, int64_t query_pitch,
                    const int32_t* probe_lists, int nprobe, int metric, int k, const uint32_t* threshold_keys,
                    int64_t* out_ids, float* out_distances, int32_t* out_cnt, int32_t* out_overflow,
                    void* workspace, size_t workspace_bytes, void* stream);
int qk_host_beta_table(int d, double* table);
int qk_aps_boundary_distances(const float* queries, int64_t num_queries, int64_t query_pitch, int d,
                              const float* centroids, int64_t centroid_pitch, const int64_t* cand_rows, int m,
                              int metric, float* out_boundary, void* stream);
int qk_aps_advance(const int32_t* active, int64_t num_active, int R, int p0, int m, int k, int d, int metric,
                   const int32_t* slots, const int64_t* round_ids, const float* round_distances,
                   const int32_t* round_cnt /* optional [num_active x R]: entries per list; else lists are padded with id -1 */,
                   const float* boundary, const double* beta_table, float recall_target,
                   float recompute_threshold, int use_precomputed, int64_t* run_ids, float* run_distances,
                   int32_t* run_count, float* radius, int32_t* have_probs, float* probs, int32_t* done,
                   int32_t* scanned, int32_t* still_active, void* stream);

/* ---- k-means assign / update / scatter ----------------------------------------------------
 * Replaces faiss::IndexFlat::search(k=1) inside faiss::Clustering::train and the final assignment
 * of kmeans() (src/cpp/src/clustering.cpp:65), and batched_scan_list(k=1) in
 * kmeans_refine_partitions (clustering.cpp:152-159): nearest centroid per point, ties to the lowest
 * centroid index. out_distances (optional): Euclidean distance / inner product to the chosen centroid.
 * Runs on the partition-scan kernel (centroids as a flat store, points as queries, k = 1), so the
 * argmin is decided in the reference's exact per-pair arithmetic. Asynchronous on `stream` (no host synchronisation). */
size_t qk_kmeans_assign_workspace_bytes(int64_t n, int64_t num_centroids, int d);
int qk_kmeans_assign(const float* points, int64_t n, int64_t point_pitch, int d,
                     const float* centroids, int64_t num_centroids, int64_t centroid_pitch,
                     int metric, int32_t* out_assign, float* out_distances,
                     void* workspace, size_t workspace_bytes, void* stream);
/* The same with the tensor-core filter's precision chosen by the caller: filter_terms 3 = 3xTF32 (what
 * qk_kmeans_assign uses), 2 = 2xTF32 (no a_lo term: 1.3x faster at K = 4096 .. 16384, d = 128). The assignment is the
 * same either way -- the winner is evaluated in exact arithmetic and a point whose proof fails is re-scanned
 * exhaustively -- but a looser filter sends more points to that slow re-scan on data whose nearest centroids are
 * closer together than its error bound. out_stats (optional, device int32[2], zeroed by the caller): [0] += points
 * re-scanned exactly, [1] = max candidates of one point -- the evidence the host's precision policy reads
 * (quake_b200/clustering.py: start with 2 terms, fall back to 3 for the rest of a training run once more than 0.5 % of
 * a call's points were re-scanned). Same workspace as qk_kmeans_assign. */
int qk_kmeans_assign_filtered(const float* points, int64_t n, int64_t point_pitch, int d,
                              const float* centroids, int64_t num_centroids, int64_t centroid_pitch,
                              int metric, int filter_terms, int32_t* out_assign, float* out_distances,
                              int32_t* out_stats, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces faiss compute_centroids (third_party/faiss/faiss/Clustering.cpp:123-192) and the scalar
 * accumulation loop of kmeans_refine_partitions (clustering.cpp:162-175): per-centroid sum of
 * assigned points and count. `order`/`offsets` come from qk_partition_by_assignment, so every
 * centroid sums its members in ascending point index (deterministic). out_sums [K x out_pitch]. */
int qk_kmeans_accumulate(const float* points, int64_t point_pitch, int d,
                         const int64_t* order, const int64_t* offsets, int64_t num_centroids,
                         float* out_sums, int64_t out_pitch, void* stream);

/* Replaces torch::sort(assignments) + bincount + split (clustering.cpp:68-84): counting sort of
 * point indices by assignment; within a centroid, ascending point index.
 * counts [K] int64, offsets [K+1] int64, order [n] int64. */
size_t qk_partition_workspace_bytes(int64_t n, int64_t num_centroids);
int qk_partition_by_assignment(const int32_t* assign, int64_t n, int64_t num_centroids,
                               int64_t* out_counts, int64_t* out_offsets, int64_t* out_order,
                               void* workspace, size_t workspace_bytes, void* stream);

/* dst[i] = src[order[i]] for rows (and ids when both id pointers are non-NULL): builds the CSR
 * arena from the sorted order (replaces index_select + per-list add_entries,
 * clustering.cpp:70-71, partition_manager.cpp:106-111). */
int qk_gather_rows(const float* src, int64_t src_pitch, const int64_t* src_ids,
                   const int64_t* order, int64_t n, int d,
                   float* dst, int64_t dst_pitch, int64_t* dst_ids, void* stream);

/* rows[i] /= ||rows[i]||_2 (in place). Replaces vectors / vectors.norm(2,1) (clustering.cpp:25-26). */
int qk_normalize_rows(float* rows, int64_t n, int64_t pitch, int d, void* stream);

/* dst[dst_rows[i]] = src[order ? order[i] : i] (rows, and ids when dst_ids != NULL): appends a batch
 * into the lists of the arena. Replaces the per-vector add_entries loop of PartitionManager::add
 * (src/cpp/src/partition_manager.cpp:245-258). */
int qk_scatter_rows(const float* src, int64_t src_pitch, const int64_t* src_ids,
                    const int64_t* order, const int64_t* dst_rows, int64_t n, int d,
                    float* dst, int64_t dst_pitch, int64_t* dst_ids, void* stream);

/* ---- device-resident bookkeeping of the dynamic partition store -------------------------------------
 * id -> arena row hash table (open addressing, power-of-two capacity from qk_hash_capacity, keys/vals are device
 * int64 arrays) and removal by id. Replaces the reference's per-list linear searches and the removal loop that walks
 * every list (src/cpp/src/dynamic_inverted_list.cpp:137-149, 302-321; src/cpp/src/index_partition.cpp:79-98, 129-145;
 * src/cpp/src/partition_manager.cpp:264-320).
 * qk_store_remove: erase ids from the table; out_rows[i] = the row the id lived in (-1: absent, or a duplicate of an
 * id already erased by this call), row_flags[row] = 1, *n_erased += rows flagged.
 * qk_store_compact_lists: one CTA per list; lists with flagged rows are compacted with the reference's
 * swap-with-last rule (the holes below the new size take, in ascending order, the surviving tail rows in descending
 * order -- the state IndexPartition::remove leaves behind, dynamic_inverted_list.cpp:127-133), the table follows the
 * moved ids, flags are cleared, new_size[list] = size after the removal. scratch_*: int32 arrays of arena length. */
int64_t qk_hash_capacity(int64_t n);
int qk_hash_clear(int64_t* keys, int64_t capacity, void* stream);
int qk_hash_insert(int64_t* keys, int64_t* vals, int64_t capacity, const int64_t* ids, const int64_t* rows,
                   int64_t n, int32_t* failed_flag, void* stream);
int qk_hash_lookup(const int64_t* keys, const int64_t* vals, int64_t capacity, const int64_t* ids, int64_t n,
                   int64_t* out_rows, void* stream);
int qk_store_remove(int64_t* keys, const int64_t* vals, int64_t capacity, const int64_t* ids, int64_t n,
                    int64_t* out_rows, uint8_t* row_flags, unsigned long long* n_erased, void* stream);
int qk_store_compact_lists(float* vectors, int64_t pitch, int64_t* ids, float* norms,
                           const int64_t* list_row0, const int64_t* list_size, int64_t num_lists,
                           int64_t* new_size, uint8_t* row_flags, int32_t* scratch_holes, int32_t* scratch_surv,
                           int64_t* hash_keys, int64_t* hash_vals, int64_t hash_capacity, void* stream);

/* ---- host-side helpers (no GPU work): vendored-faiss control logic of Clustering::train ----------
 * First m entries of faiss::rand_perm(n, seed) (third_party/faiss/faiss/utils/random.cpp:153-163). */
int qk_host_rand_perm_prefix(int64_t n, int64_t seed, int64_t m, int64_t* out);
/* faiss split_clusters (third_party/faiss/faiss/Clustering.cpp:204-251) on HOST arrays. */
int qk_host_split_clusters(int64_t d, int64_t k, int64_t n, float* hassign, float* centroids,
                           int64_t pitch, int64_t* nsplit_out);

#ifdef __cplusplus
}
#endif
#endif /* QUAKE_B200_H */
