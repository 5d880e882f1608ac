// Internal structures of libflatnav_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "fnb_layout.h"
#include "search_kernel.cuh"

#define FNB_HEADER_BYTES 60u

namespace fnb {

struct Header {
  int32_t data_type = 0;
  int32_t metric = 0;
  uint64_t M = 0, data_size = 0, node_size = 0, max_nodes = 0, cur_nodes = 0, dim = 0;
};

// One full copy of the index in the HBM of one device, plus the per-device launch state.
struct Replica {
  int device = -1;
  int num_sms = 0;
  uint4* vec = nullptr;
  uint32_t* adj = nullptr;
  int32_t* labels = nullptr;
  unsigned int* counter = nullptr;
  unsigned long long* totals = nullptr;
  unsigned long long* h_totals = nullptr;  // pinned mirror of `totals`
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  unsigned char* ws = nullptr;  // device workspace of the host-buffer entry points
  size_t ws_bytes = 0;
  unsigned char* h_pinned = nullptr;      // pinned staging block of the small-batch path
  unsigned char* h_pinned_dev = nullptr;  // its device-side alias
  size_t h_pinned_bytes = 0;
  uint64_t device_bytes = 0;
  uint64_t capacity = 0;  // rows the vec / adj / labels arrays can hold (>= cur_nodes; construction grows it)
};

int fail(int code, const char* fmt, ...);
extern thread_local std::string g_last_error;

}  // namespace fnb

struct fnb_index {
  fnb::Header h;
  uint32_t nchunks = 0, stride = 0;
  int G = 8;
  std::vector<fnb::Replica> replicas;
  std::mutex mu;
};

// build.cu: host vectors -> padded rows [cur_nodes, cur_nodes + n), labels (NULL: label_base, label_base + 1, ...),
// optionally all-self-loop link rows.  cur_nodes is left to the caller, who holds index->mu.
int upload_new_rows_locked(fnb_index* ix, const void* vectors, const int32_t* labels, int32_t label_base, int64_t n,
                           bool init_links);

namespace fnb {
int plan_search(const fnb_index* ix, int64_t Q, int K, int ef, int ninit, SearchParams* p, int64_t launch_q = 0);
cudaError_t dispatch_search(const fnb_index* ix, const SearchParams& p, int num_sms, cudaStream_t s);
cudaError_t dispatch_search_f32(const fnb_index* ix, const SearchParams& p, int num_sms, cudaStream_t s);
cudaError_t dispatch_search_u8(const fnb_index* ix, const SearchParams& p, int num_sms, cudaStream_t s);
cudaError_t dispatch_search_i8(const fnb_index* ix, const SearchParams& p, int num_sms, cudaStream_t s);

template <int DT, int METRIC>
static inline cudaError_t dispatch_gc(const fnb_index* ix, const SearchParams& p, int num_sms, cudaStream_t s) {
  const int ch = fnb_chunks_per_lane(ix->nchunks, ix->G);
  if (ix->G == 4) {
    if (ch <= 1) return launch_search<DT, METRIC, 4, 1>(p, num_sms, s);
    return launch_search<DT, METRIC, 4, 2>(p, num_sms, s);
  }
  if (ix->G == 8) {
    switch (ch) {
      case 1: return launch_search<DT, METRIC, 8, 1>(p, num_sms, s);
      case 2: return launch_search<DT, METRIC, 8, 2>(p, num_sms, s);
      case 3: return launch_search<DT, METRIC, 8, 3>(p, num_sms, s);
      default: return launch_search<DT, METRIC, 8, 4>(p, num_sms, s);
    }
  }
  if (ch <= 2) return launch_search<DT, METRIC, 32, 2>(p, num_sms, s);
  if (ch <= 4) return launch_search<DT, METRIC, 32, 4>(p, num_sms, s);
  if (ch <= 8) return launch_search<DT, METRIC, 32, 8>(p, num_sms, s);
  return launch_search<DT, METRIC, 32, 16>(p, num_sms, s);
}
}  // namespace fnb
