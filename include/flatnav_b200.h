/*
 * flatnav_b200 — C ABI of the B200-native batched query engine for FlatNav's search hot path.
 *
 * The reference (BlaiseMuhirwa/flatnav) has no C ABI: its boundary for this path is the C++
 * template `flatnav::Index<dist_t,label_t>` and the pybind11 class built on it.  This header is the
 * thin, language-neutral layer placed beneath both surfaces; every entry point cites the reference
 * interface it replaces (paths relative to the reference checkout).  The C++ shim
 * (include/flatnav_b200/Index.h) and the Python package (flatnav_b200/) are written on top of it,
 * and INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions: every function returns FNB_OK (0), a positive informational status
 * (FNB_SHORT_RESULT) or a negative error code, and never throws across the boundary;
 * fnb_last_error() returns a thread-local message for the last non-zero status.  All pointers are
 * plain host or device pointers owned by the caller unless stated otherwise; no torch types appear.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * FNB_ERR_CUDA.
 */
#ifndef FLATNAV_B200_H_
#define FLATNAV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* flatnav::util::DataType values that may appear in an index file (include/flatnav/util/Datatype.h:11-24);
 * the enum value is part of the on-disk format (serialised as int32). */
enum { FNB_DTYPE_UINT8 = 0, FNB_DTYPE_INT8 = 4, FNB_DTYPE_FLOAT32 = 9, FNB_DTYPE_ANY = -1 };

/* flatnav::distances::MetricType (DistanceInterface.h).  The metric is NOT stored in the file — in the
 * reference it is implied by the C++ class used to load (SquaredL2Distance / InnerProductDistance) — so
 * the caller states it. */
enum { FNB_METRIC_L2 = 0, FNB_METRIC_IP = 1 };

enum {
  FNB_OK = 0,
  FNB_SHORT_RESULT = 1,      /* some query reached fewer than K nodes: slots hold label -1, distance +inf.
                                (bindings.cpp:134-137,184-189 raise RuntimeError for this) */
  FNB_ERR_INVALID_ARG = -1,  /* std::invalid_argument in the reference (e.g. num_initializations <= 0, Index.h:847) */
  FNB_ERR_IO = -2,           /* std::runtime_error "Unable to open file" (Index.h:445-447, 484-486) */
  FNB_ERR_FORMAT = -3,       /* header inconsistent / file truncated / dtype or metric mismatch */
  FNB_ERR_CUDA = -4,         /* CUDA runtime failure or no device */
  FNB_ERR_UNSUPPORTED = -5,  /* dimension / ef beyond what the kernels are instantiated for */
  FNB_ERR_NOMEM = -6
};

typedef struct fnb_index fnb_index; /* opaque; owns device memory on every device it was loaded on */

/* Mirrors the getters of Index<> (Index.h:517-531) plus what the loader decided. */
typedef struct fnb_info {
  int32_t data_type;          /* Index::getDataType */
  int32_t metric;             /* FNB_METRIC_* as given at load time */
  uint64_t max_edges_per_node; /* Index::maxEdgesPerNode (M) */
  uint64_t dim;               /* Index::dataDimension */
  uint64_t data_size_bytes;   /* Index::dataSizeBytes */
  uint64_t node_size_bytes;   /* Index::nodeSizeBytes = data_size + 4 M + 4 */
  uint64_t max_node_count;    /* Index::maxNodeCount */
  uint64_t cur_num_nodes;     /* Index::currentNumNodes */
  int32_t n_devices;          /* replicas (shard_mode 0) */
  int32_t device_ids[16];
  uint64_t device_bytes;      /* HBM bytes held per device */
  uint32_t row_stride_bytes;  /* padded vector row pitch in HBM */
  uint32_t lanes_per_row;     /* G of the distance kernel (csrc/fnb_layout.h) */
} fnb_info;

/* Counters of one fnb_search call (sums over the batch); feed the roofline accounting of bench.py. */
typedef struct fnb_search_stats {
  int64_t n_queries;
  int64_t n_dist;        /* database-row distance evaluations incl. entry-selection probes */
  int64_t n_hops;        /* expanded nodes */
  int64_t n_short;       /* queries with fewer than K results */
  int64_t algo_bytes;    /* n_dist*D*s + n_hops*M*4 + Q*D*s + Q*K*8    */
  float kernel_ms;       /* device time of the traversal kernel(s), max over devices (CUDA events).  Recorded when the
                            call synchronises its stream anyway, or when FNB_TIME_KERNELS=1 is set in the environment;
                            0 otherwise (a call that completes by the pinned-memory flag records no events) */
  float total_ms;        /* device time incl. host<->device copies (host-buffer entry point only); as kernel_ms */
  int32_t kernel_launches;
  int32_t reserved;
} fnb_search_stats;

/* ---- load / save ----------------------------------------------------------------------------- */

/* Replaces Index<dist_t,label_t>::loadIndex(filename) (include/flatnav/index/Index.h:442-479) and
 * PyIndex::loadIndex (python-bindings/src/flatnav/bindings.cpp:303-306).  Parses the 60-byte cereal
 * header + node blob, validates it, and lays the index out in HBM (vectors / links / labels as
 * separate arrays) on each of `n_devices` devices (a full replica per device).
 * expect_dtype: FNB_DTYPE_* or FNB_DTYPE_ANY.  device_ids may be NULL => {current device}. */
int fnb_index_load(const char* path, int metric, int expect_dtype, const int* device_ids, int n_devices,
                   fnb_index** out);

/* Same, from an in-memory image of the file (header + blob). */
int fnb_index_from_memory(const void* file_bytes, size_t nbytes, int metric, int expect_dtype, const int* device_ids,
                          int n_devices, fnb_index** out);

/* Replaces Index::saveIndex (Index.h:481-490): writes the reference's byte layout back (round trip). */
int fnb_index_save(const fnb_index* index, const char* path);

int fnb_index_info(const fnb_index* index, fnb_info* out);
void fnb_index_free(fnb_index* index);

/* ---- construction (the caller side of the search path) --------------------------------- */

/* Replaces the Index constructor (include/flatnav/index/Index.h:159-180) and flatnav.index.create
 * (python-bindings/src/flatnav/bindings.cpp:484-504): an empty index for up to max_node_count vectors on `device`
 * (-1 = current device). */
int fnb_index_create(int metric, int data_type, uint64_t dim, uint64_t max_node_count, uint64_t max_edges_per_node,
                     int device, fnb_index** out);

/* Grows the device arrays of a single-device index so that fnb_index_add can append up to max_node_count nodes
 * (a loaded index holds exactly cur_num_nodes rows). */
int fnb_index_reserve(fnb_index* index, uint64_t max_node_count);

typedef struct fnb_build_stats {
  int64_t n_added;
  int64_t n_batches;            /* insertion batches (each: search, select + link, prune) */
  int64_t n_dropped_backlinks;  /* back-links not considered because one node got > 96 newcomers in one batch */
  float device_ms;              /* upload + construction, CUDA events */
  float reserved;
} fnb_build_stats;

/* Replaces Index::addBatch / add (Index.h:301-378) with selectNeighbors (:714-763) and connectNeighbors (:765-834),
 * and PyIndex::add (bindings.cpp:62-110): appends n vectors (HOST, row-major [n, dim] of the index data type) with
 * their labels (HOST int32 [n]; NULL = 0 .. n-1 like the binding) and links them into the graph.  Insertion is
 * batched on the GPU: the traversal kernel finds each new node's ef_construction nearest inserted nodes, then the
 * reference's pruning heuristic picks max(M/2, 1) links and the back-links are merged (a full row is re-pruned to
 * M).  The graph is not node-for-node the reference's (neither are two multi-threaded reference builds); its
 * search quality is, and fnb_index_save writes it in the reference's format. */
int fnb_index_add(fnb_index* index, const void* vectors, const int32_t* labels, int64_t n, int ef_construction,
                  int num_initializations, fnb_build_stats* stats);

/* ---- link import and graph re-ordering ------------------------------------ */

/* Replaces PyIndex::allocateNodes / Index::allocateNode (bindings.cpp:308-324, Index.h:262-272): appends n vectors
 * (HOST, row-major [n, dim] of the index data type) as unlinked nodes — every link slot a self-loop.  labels: HOST
 * int32 [n], or NULL for cur_num_nodes, cur_num_nodes + 1, ... (the binding numbers them with a running counter). */
int fnb_index_allocate_nodes(fnb_index* index, const void* vectors, const int32_t* labels, int64_t n);

/* Replaces Index::buildGraphLinks (Index.h:187-238): reads a Matrix Market edge list (1-based "u v" lines after the
 * '%' comments and a size line "rows cols M"); edge (u, v) takes the first slot of u's link row that still points at
 * u, in file order; edges beyond a full row are dropped.  Same checks and messages as the reference: rows must equal
 * max_node_count and the third number max_edges_per_node (FNB_ERR_IO, a std::runtime_error there). */
int fnb_index_build_graph_links(fnb_index* index, const char* mtx_filename);

/* Replaces Index::getGraphOutdegreeTable (Index.h:240-251) at the ABI level: copies the link table to HOST memory,
 * uint32 [cur_num_nodes, max_edges_per_node], self-loops (unused slots) included — the caller drops them. */
int fnb_index_links(const fnb_index* index, uint32_t* out_links);

/* Replaces Index::reorderGOrder(window) / reorderRCM / one step of doGraphReordering (Index.h:412-440) with
 * util::gOrder / util::rcmOrder (util/Reordering.h:26-199) and Index::relabel (Index.h:872-926).  The ordering is
 * computed on the host from the link table (sequential greedy algorithms; same queue discipline as the reference, so
 * the permutation — and the file fnb_index_save then writes — is the reference's, byte for byte); the relabelling
 * of all links and the re-layout of vectors / link rows / labels run on the GPU on every replica.
 * window: gorder window size, <= 0 selects the reference's default 5 (ignored for RCM).
 * perm_out: optional HOST uint32 [cur_num_nodes], perm_out[old node id] = new node id. */
enum { FNB_REORDER_GORDER = 0, FNB_REORDER_RCM = 1 };
int fnb_index_reorder(fnb_index* index, int method, int window, uint32_t* perm_out);

/* The ordering step of fnb_index_reorder alone, on a HOST link table uint32 [n_nodes, max_edges_per_node] (self-loops
 * = unused slots): util::gOrder / util::rcmOrder (util/Reordering.h:26-199).  Host-only graph algorithm, needs no
 * device; perm_out[old node id] = new node id. */
int fnb_graph_order(const uint32_t* links, uint64_t n_nodes, uint64_t max_edges_per_node, int method, int window,
                    uint32_t* perm_out);

/* Index::relabel (Index.h:872-926) with a caller-supplied permutation (HOST uint32 [cur_num_nodes],
 * perm[old node id] = new node id): node i moves to row perm[i], every link is mapped through perm.  Search results
 * are unchanged up to the order of exact distance ties (labels travel with their nodes).  FNB_ERR_INVALID_ARG if
 * perm is not a permutation. */
int fnb_index_relabel(fnb_index* index, const uint32_t* perm);

/* ---- search ------------------------------------------------------------------------------------ */

/* Replaces the batched fan-out PyIndex::searchImpl (bindings.cpp:161-228: executeInParallel over
 * Index::search, Index.h:387-409) and, with Q == 1, Index::search / PyIndex::searchSingleImpl
 * (bindings.cpp:121-159).  HOST buffers: `queries` is row-major [Q, dim] of the index data type,
 * out_dist float32 [Q, K], out_label int32 [Q, K]; copies to/from the device happen inside.
 * With several replicas the queries are split evenly across devices.
 * Results per query: ascending distance, squared-L2 or 1 - <x,y> exactly as the reference returns them;
 * labels are the node label field (Index.h:396-399).  stats may be NULL.  Thread-safe per index. */
int fnb_search(fnb_index* index, const void* queries, int64_t Q, int K, int ef_search, int num_initializations,
               float* out_dist, int32_t* out_label, fnb_search_stats* stats);

/* Same traversal with DEVICE-resident buffers on replica `replica` (0-based), enqueued on `cuda_stream`
 * (a cudaStream_t; NULL = legacy default stream) and not synchronised: the kernel-only path.
 * d_ndist / d_nhops: optional per-query uint32 counters in device memory (may be NULL). */
int fnb_search_device(fnb_index* index, int replica, const void* d_queries, int64_t Q, int K, int ef_search,
                      int num_initializations, float* d_out_dist, int32_t* d_out_label, uint32_t* d_ndist,
                      uint32_t* d_nhops, void* cuda_stream);

/* Measurement aid (no reference equivalent): the name of the traversal-kernel template instantiation a search of this
 * shape launches on this index, spelled the way ncu / cuobjdump print it without blanks, e.g.
 * "fnb_search_kernel<0,0,8,4,1,0,0>" (data type, metric, lanes per row, chunks per lane, exact fit, latency variant,
 * occupancy plan).  bench.py uses it to check that a committed ncu capture is of the kernel it times. */
int fnb_search_kernel_signature(const fnb_index* index, int64_t Q, int K, int ef_search, char* out, size_t out_capacity);

/* Measurement aid: the launch plan of a search of this shape (what bounds the resident queries per SM). */
typedef struct fnb_plan_info {
  int32_t latency_variant;      /* batches of at most 4 x SMs queries: 2 = one CTA of four warps per query (default),
                                   1 = one warp per CTA; 0 = throughput kernel (one warp per query, 4 per CTA) */
  int32_t dense_plan;           /* 1: the 28-warps-per-SM instantiation (large batches of rows up to 512 B) */
  int32_t list_capacity;        /* max(ef_search, K) rounded up to 32 entries of 8 bytes */
  int32_t visited_slots;        /* tag slots of the per-query visited set */
  int32_t smem_bytes_per_query; /* list + visited set + scratch */
  int32_t ctas_per_sm;          /* planned resident CTAs per SM (min of the register plan and what shared memory allows) */
  int32_t warps_per_sm;
  int32_t queries_per_sm;       /* resident queries per SM */
} fnb_plan_info;
int fnb_search_plan(const fnb_index* index, int64_t Q, int K, int ef_search, fnb_plan_info* out);

/* Exact scan over all nodes (ground truth / exact re-rank).  The reference has no brute force of its own;
 * semantics: top-K by (distance, node id), distances in the same arithmetic as fnb_search.  HOST buffers. */
int fnb_bruteforce(fnb_index* index, const void* queries, int64_t Q, int K, float* out_dist, int32_t* out_label);

/* Exact re-rank of caller-supplied candidates (the reference's benchmark driver,
 * experiments/run-benchmark.py:38-124, has no re-rank step of its own — extension).  For each of the Q queries (HOST,
 * row-major [Q, dim] of the index data type) the C entries of candidates[q] (HOST int32 [Q, C]) are evaluated with the
 * arithmetic of fnb_search — the distance reported for a (query, node) pair is bit-identical across fnb_search,
 * fnb_bruteforce and fnb_rerank — and the K best by (distance, node id) are returned with their label fields
 * (float32 / int32 [Q, K]; unfilled slots +inf / -1).  candidates_are_labels != 0: entries are node labels as search
 * returns them (unknown labels and negative entries are skipped, of equal labels the lowest node id answers);
 * 0: entries are node ids.  A candidate listed twice counts once. */
int fnb_rerank(fnb_index* index, const void* queries, int64_t Q, const int32_t* candidates, int C,
               int candidates_are_labels, int K, float* out_dist, int32_t* out_label);

/* What the last fnb_bruteforce call on this thread did.  Large problems run as a tcgen05 (tensor-core) GEMM
 * over bf16 hi/lo splits of the vectors that only FILTERS candidates, followed by an exact re-rank in the
 * search kernel's arithmetic; a query whose candidate list cannot be proven complete is re-scanned exactly, so
 * the result is always the exact top-K by (distance, node id).  No reference equivalent (extension). */
typedef struct fnb_bf_stats {
  int32_t path;          /* 0 = CUDA-core exact scan, 1 = tensor-core filter + exact re-rank */
  int32_t reserved;
  int64_t n_unsafe;      /* queries re-scanned exactly because the filter margin could not be proven */
  int64_t n_candidates;  /* candidates re-ranked exactly (sum over queries) */
  float prep_ms, gemm_ms, rerank_ms, rescan_ms; /* device times of the four phases */
  double gemm_flops;     /* 2*Q*N*Dpad*passes issued to the tensor cores */
} fnb_bf_stats;
int fnb_bruteforce_stats(fnb_bf_stats* out);

/* Dataset-sharded search: k-way merge of `n_lists` per-shard result lists (each [Q, K], ascending) that
 * were gathered into DEVICE buffers d_dist / d_label of shape [n_lists, Q, K]; writes the global top-K
 * ([Q, K], ties -> lower label) to d_out_*.  Enqueued on cuda_stream. */
int fnb_merge_topk(const float* d_dist, const int32_t* d_label, int n_lists, int64_t Q, int K, float* d_out_dist,
                   int32_t* d_out_label, void* cuda_stream);

/* ---- dataset-sharded search over NVLink peer memory (one process per GPU) ------------------------------ */
/* The reference has no multi-device mode (its only parallelism is executeInParallel over queries,
 * include/flatnav/util/Multithreading.h:18-48).  Every rank holds one sub-graph (labels = global ids), answers all
 * queries on it, and the per-shard top-K lists are merged.  fnb_search_sharded does the exchange without a
 * collective library: after the local traversal ONE kernel pushes this rank's lists into every peer's gather
 * buffer through CUDA-IPC-mapped pointers, publishes / awaits per-rank epoch flags and runs the k-way merge
 * (ties -> lower label).  Usage: create on every rank, exchange the 64-byte handles out of band (e.g.
 * torch.distributed.all_gather_object), attach, then call fnb_search_sharded collectively (same Q, K everywhere). */
#define FNB_IPC_HANDLE_BYTES 64
typedef struct fnb_exchange fnb_exchange;
int fnb_exchange_create(int device, int rank, int world, int64_t max_Q, int max_K, fnb_exchange** out);
int fnb_exchange_handle(fnb_exchange* ex, void* handle_out /* FNB_IPC_HANDLE_BYTES */);
int fnb_exchange_attach(fnb_exchange* ex, const void* handles /* world x FNB_IPC_HANDLE_BYTES, rank order */);
/* DEVICE buffers; enqueued on cuda_stream, not synchronised.  d_out_* receive the global top-K [Q, K]. */
int fnb_search_sharded(fnb_index* shard, fnb_exchange* ex, const void* d_queries, int64_t Q, int K, int ef_search,
                       int num_initializations, float* d_out_dist, int32_t* d_out_label, void* cuda_stream);
int fnb_exchange_status(fnb_exchange* ex); /* after a synchronise: FNB_ERR_CUDA if a peer never delivered */
void fnb_exchange_free(fnb_exchange* ex);

const char* fnb_last_error(void);
const char* fnb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FLATNAV_B200_H_ */
