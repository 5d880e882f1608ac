// Internal structures of libflatnav_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <vector>

#include "fnb_layout.h"
#include "search_cta_kernel.cuh"
#include "search_kernel.cuh"

#define FNB_HEADER_BYTES 60u

namespace fnb {

struct Header {
  int32_t data_type = 0;
  int32_t metric = 0;
  uint64_t M = 0, data_size = 0, node_size = 0, max_nodes = 0, cur_nodes = 0, dim = 0;
};

// One search in flight: everything a host-buffer search call needs besides the index arrays.  A replica keeps a pool of
// lanes so that concurrent callers of fnb_search (the reference's Index::search is re-entrant, Index.h:387-409 with
// VisitedSetPool.h:154-172 handing every thread its own visited set) run side by side on their own streams.
struct Lane {
  cudaStream_t stream = nullptr;
  unsigned int* counter = nullptr;         // device: persistent-warp work counter
  unsigned long long* totals = nullptr;    // device: [3] n_dist, n_hops, short results
  unsigned int* q_ready = nullptr;         // device: watermark of a host-fed batch (queries already copied in)
  unsigned long long* h_totals = nullptr;  // pinned: the kernel's last warp writes the totals of the launch here
  unsigned long long* h_totals_dev = nullptr;  // its device-side alias
  volatile uint32_t* h_flag = nullptr;     // pinned: sequence number of the lane's last launch that has finished
  unsigned int* h_flag_dev = nullptr;      // its device-side alias
  uint32_t seq = 0;                        // launches made on this lane
  uint32_t* h_marks = nullptr;             // pinned: the watermark values a host-fed batch copies to q_ready, [FNB_FEED_CHUNKS + 1]
  cudaStream_t copy_stream = nullptr;      // feeds the queries of a pageable caller while the kernel runs on `stream`
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_feed = nullptr;
  unsigned char* ws = nullptr;  // device workspace: queries of pageable callers, results too large for the pinned block
  size_t ws_bytes = 0;
  unsigned char* h_pinned = nullptr;      // pinned block the kernel writes results / per-query counters into
  unsigned char* h_pinned_dev = nullptr;  // its device-side alias
  size_t h_pinned_bytes = 0;
};

#define FNB_FEED_CHUNKS 48
#define FNB_RING_SLOTS 256u  // per-call launch-state slots of fnb_search_device (64 B each)
// Layout of a 64-byte launch-state slot (ring slots, and the first half of a lane's counter block).  A slot is all
// zero between launches: the last warp of a launch leaves it so (search_kernel.cuh, `done`).
#define FNB_SLOT_COUNTER 0      /* u32  persistent-warp work counter */
#define FNB_SLOT_DONE 4         /* u32  warps that have left the kernel */
#define FNB_SLOT_TOTALS 8       /* u64 x 3  n_dist, n_hops, short results (accumulated with atomics) */
#define FNB_SLOT_LAST_TOTALS 32 /* u64 x 3  the totals of the launch that used the slot last (read by the host) */

// One caller's share of a combined latency launch.
struct CombReq {
  const unsigned char* q = nullptr;
  int64_t nq = 0;
  int K = 0, ef = 0, ninit = 0;
  float* out_dist = nullptr;
  int32_t* out_label = nullptr;
  int64_t nd = 0, nh = 0, ns = 0;
  int rc = 0;
  std::string err;
  bool taken = false, done = false;
};

struct LanePool {
  std::mutex mu;
  std::condition_variable cv;
  std::vector<Lane*> all, idle;
  int max_lanes = 16;
  unsigned char* ring = nullptr;  // device: FNB_RING_SLOTS x 64 B
  std::atomic<uint32_t> ring_seq{0};
  volatile unsigned int* h_done_seq = nullptr;  // pinned: sequence number of the last fnb_search_device launch that finished
  unsigned int* d_done_seq = nullptr;           // its device-side alias
  // Combining of concurrent latency batches (fnb_search from many host threads): requests queue here; whoever finds
  // no launch being prepared takes every compatible request waiting and launches them as ONE kernel.
  std::mutex cmu;
  std::condition_variable ccv;
  std::vector<struct CombReq*> cpending;
  bool claunching = false;
};

// One full copy of the index in the HBM of one device, plus the per-device launch state.
struct Replica {
  int device = -1;
  int num_sms = 0;
  uint4* vec = nullptr;
  uint32_t* adj = nullptr;
  int32_t* labels = nullptr;
  // stream / counters of the mutating entry points (construction, re-ordering, brute force): used under the
  // exclusive index lock only
  unsigned int* counter = nullptr;
  unsigned long long* totals = nullptr;
  cudaStream_t stream = nullptr;
  LanePool* pool = nullptr;  // search lanes (shared index lock)
  uint64_t device_bytes = 0;
  uint64_t capacity = 0;  // rows the vec / adj / labels arrays can hold (>= cur_nodes; construction grows it)
};

int fail(int code, const char* fmt, ...);
extern thread_local std::string g_last_error;

}  // namespace fnb

struct fnb_build_scratch;  // build.cu
void fnb_build_scratch_free(fnb_build_scratch* b);
struct fnb_label_map;  // rerank.cu
void fnb_label_map_free(fnb_label_map* m);
struct fnb_index;
// Caches derived from the graph (construction degrees, label table) are dropped by whoever edits it; the caller holds
// the exclusive index lock.  links: link rows changed outside fnb_index_add; labels: labels / node numbering changed.
void fnb_index_mutated(fnb_index* ix, bool links, bool labels);

struct fnb_index {
  fnb::Header h;
  uint32_t nchunks = 0, stride = 0;
  int G = 8;
  std::vector<fnb::Replica> replicas;
  fnb_build_scratch* build = nullptr;  // construction scratch kept between fnb_index_add calls (single-device index)
  fnb_label_map* label_map = nullptr;  // label -> node table of fnb_rerank, built on first use
  std::mutex aux_mu;                   // guards label_map creation among concurrent (shared-lock) callers
  // searches (fnb_search, fnb_search_device, save, info) hold it shared; everything that changes the graph or the
  // device arrays (add, reserve, allocate_nodes, build_graph_links, relabel, reorder) and brute force hold it
  // exclusively and first wait for every kernel still in flight on the index's devices (fnb::quiesce)
  mutable std::shared_mutex mu;
};

// build.cu: host vectors -> padded rows [cur_nodes, cur_nodes + n), labels (NULL: label_base, label_base + 1, ...),
// optionally all-self-loop link rows.  cur_nodes is left to the caller, who holds index->mu.
int upload_new_rows_locked(fnb_index* ix, const void* vectors, const int32_t* labels, int32_t label_base, int64_t n,
                           bool init_links);

namespace fnb {
typedef std::unique_lock<std::shared_mutex> ExclusiveLock;
typedef std::shared_lock<std::shared_mutex> SharedLock;
// after taking the exclusive lock: wait for asynchronous searches (fnb_search_device on caller streams) still in flight
void quiesce(fnb_index* ix);
// n_nodes != 0: search only the first n_nodes nodes (construction: the graph a batch is inserted into)
// allow_latency_variant = false: always the throughput kernel (construction)
int plan_search(const fnb_index* ix, int64_t Q, int K, int ef, int ninit, SearchParams* p, int64_t launch_q = 0,
                uint64_t n_nodes = 0, bool allow_latency_variant = true);
cudaError_t dispatch_search(const fnb_index* ix, const SearchParams& p, int num_sms, cudaStream_t s);
cudaError_t dispatch_search_f32(const fnb_index* ix, const SearchParams& p, int num_sms, cudaStream_t s);
cudaError_t dispatch_search_u8(const fnb_index* ix, const SearchParams& p, int num_sms, cudaStream_t s);
cudaError_t dispatch_search_i8(const fnb_index* ix, const SearchParams& p, int num_sms, cudaStream_t s);

template <int DT, int METRIC, int G, int CH>
static inline cudaError_t launch_any(const SearchParams& p, int num_sms, cudaStream_t s) {
  return p.lat == 2u ? launch_search_cta<DT, METRIC, G, CH>(p, num_sms, s) : launch_search<DT, METRIC, G, CH>(p, num_sms, s);
}

template <int DT, int METRIC>
static inline cudaError_t dispatch_gc(const fnb_index* ix, const SearchParams& p, int num_sms, cudaStream_t s) {
  const int ch = fnb_chunks_per_lane(ix->nchunks, ix->G);
  if (ix->G == 4) {
    if (ch <= 1) return launch_any<DT, METRIC, 4, 1>(p, num_sms, s);
    return launch_any<DT, METRIC, 4, 2>(p, num_sms, s);
  }
  if (ix->G == 8) {
    switch (ch) {
      case 1: return launch_any<DT, METRIC, 8, 1>(p, num_sms, s);
      case 2: return launch_any<DT, METRIC, 8, 2>(p, num_sms, s);
      case 3: return launch_any<DT, METRIC, 8, 3>(p, num_sms, s);
      default: return launch_any<DT, METRIC, 8, 4>(p, num_sms, s);
    }
  }
  if (ch <= 2) return launch_any<DT, METRIC, 32, 2>(p, num_sms, s);
  if (ch <= 4) return launch_any<DT, METRIC, 32, 4>(p, num_sms, s);
  if (ch <= 8) return launch_any<DT, METRIC, 32, 8>(p, num_sms, s);
  return launch_any<DT, METRIC, 32, 16>(p, num_sms, s);
}
}  // namespace fnb
