// Dataset-sharded search, exchange step: k-way merge of per-shard top-K lists into the global top-K.
// Each shard is an ordinary flatnav index over a contiguous id range whose labels are global ids
// (SURVEY.md §8e); after the per-shard [Q,K] results have been all-gathered into one buffer
// [n_lists, Q, K], one warp per query merges them: lane l owns the head of list l, a warp-wide
// min picks the next output.  Ties -> lower label.
#include "../../include/flatnav_b200.h"
#include "fnb_internal.h"

namespace fnb {

__global__ void merge_topk_kernel(const float* __restrict__ d_dist, const int32_t* __restrict__ d_label, int n_lists,
                                  uint32_t Q, uint32_t K, float* __restrict__ out_dist, int32_t* __restrict__ out_label) {
  const int lane = threadIdx.x & 31;
  const uint32_t qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (qi >= Q) return;
  const size_t plane = (size_t)Q * K;
  const size_t base = (size_t)lane * plane + (size_t)qi * K;
  uint32_t head = 0;
  auto load = [&](uint32_t h) -> uint64_t {
    if (lane >= n_lists || h >= K) return ~0ull;
    const int32_t lab = d_label[base + h];
    if (lab < 0) return ~0ull;  // unfilled slot of a short per-shard result
    return ((uint64_t)ord_f32(d_dist[base + h]) << 32) | (uint32_t)lab;
  };
  uint64_t cur = load(0);
  for (uint32_t j = 0; j < K; j++) {
    uint64_t best = cur;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const uint64_t o = shfl64(best, lane ^ off);
      best = o < best ? o : best;
    }
    if (lane == 0) {
      out_dist[(size_t)qi * K + j] = best == ~0ull ? __int_as_float(0x7f800000) : unord_f32((uint32_t)(best >> 32));
      out_label[(size_t)qi * K + j] = best == ~0ull ? -1 : (int32_t)(uint32_t)best;
    }
    if (best != ~0ull && cur == best) {  // labels are unique across shards, so exactly one lane advances
      head++;
      cur = load(head);
    }
  }
}

}  // namespace fnb

using namespace fnb;

extern "C" int fnb_merge_topk(const float* d_dist, const int32_t* d_label, int n_lists, int64_t Q, int K,
                              float* d_out_dist, int32_t* d_out_label, void* cuda_stream) {
  if (n_lists < 1 || n_lists > 32) return fail(FNB_ERR_INVALID_ARG, "n_lists must be in [1, 32]");
  if (Q < 0 || K <= 0) return fail(FNB_ERR_INVALID_ARG, "bad Q or K");
  if (Q == 0) return FNB_OK;
  if (!d_dist || !d_label || !d_out_dist || !d_out_label) return fail(FNB_ERR_INVALID_ARG, "NULL buffer");
  const int threads = 128;
  const long long blocks = (Q * 32 + threads - 1) / threads;
  merge_topk_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)cuda_stream>>>(d_dist, d_label, n_lists, (uint32_t)Q,
                                                                               (uint32_t)K, d_out_dist, d_out_label);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(FNB_ERR_CUDA, "merge kernel launch failed: %s", cudaGetErrorString(e));
  return FNB_OK;
}
