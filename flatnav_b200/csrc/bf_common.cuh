// Shared pieces of the two exact-scan paths (bruteforce.cu: CUDA-core scan; bruteforce_tc.cu: tcgen05 candidate
// filter + exact re-rank).
#pragma once
#include "fnb_internal.h"

namespace fnb {

#define BF_WARPS 8
#define BF_TILE_ROWS 32

struct BfParams {
  const uint4* __restrict__ vec;
  const int32_t* __restrict__ labels;
  const void* __restrict__ queries;
  const uint32_t* __restrict__ qmap;  // optional: slot -> query row (re-scan of a subset); null = identity
  float* __restrict__ out_dist;
  int32_t* __restrict__ out_label;
  uint32_t N, dim, nchunks, stride, Q, K, Kcap, query_vec_ok;
};

// Insert `kx` = (ordered distance << 32 | node id) into the ascending list of at most K keys held in shared
// memory by one warp (all lanes call with the same kx).  Rare after the first few rows: O(len / 32) steps.
__device__ __forceinline__ void warp_topk_insert(uint64_t* list, uint32_t& len, const uint32_t K, const uint64_t kx,
                                                 const int lane) {
  if (len >= K && !(kx < list[len - 1])) return;
  uint32_t cnt = 0;
  for (uint32_t i = lane; i < len; i += 32) cnt += (list[i] < kx) ? 1u : 0u;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(FNB_FULL, cnt, off);
  const uint32_t ins = cnt;
  for (int c = (int)((len ? len - 1 : 0) >> 5); c >= (int)(ins >> 5) && len; c--) {
    const uint32_t i = (uint32_t)c * 32 + lane;
    const bool have = i < len && i >= ins;
    const uint64_t y = have ? list[i] : 0ull;
    __syncwarp();
    if (have && i + 1 < K) list[i + 1] = y;
    __syncwarp();
  }
  if (lane == 0) list[ins] = kx;
  __syncwarp();
  len = min(K, len + 1);
}

// entry points of the two paths (host)
struct BfRun {          // what the last fnb_bruteforce call did (fnb_bruteforce_stats)
  int path = 0;         // 0 = CUDA-core exact scan, 1 = tcgen05 filter + exact re-rank
  int64_t n_unsafe = 0; // queries of the tensor path that were re-scanned exactly
  int64_t n_candidates = 0;
  float prep_ms = 0.f, gemm_ms = 0.f, rerank_ms = 0.f, rescan_ms = 0.f;
  double gemm_flops = 0.0;
};
cudaError_t launch_exact_scan(const fnb_index* ix, const BfParams& p, cudaStream_t s);
int bruteforce_tensor(fnb_index* ix, Replica& r, const void* d_queries, int64_t Q, int K, float* d_out_dist,
                      int32_t* d_out_label, BfRun* run);
bool tensor_path_supported(const fnb_index* ix, int64_t Q, int K);

}  // namespace fnb
