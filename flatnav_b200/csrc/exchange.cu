// Dataset-sharded search (BASELINE.json cfg5), the exchange step over NVLink peer memory.
//
// Every rank (one process per GPU) holds ONE sub-graph whose labels are global ids, answers ALL queries on it, and
// the per-shard top-K lists are combined into the global top-K.  The reference has no multi-device mode at all (its
// only parallelism is executeInParallel over queries, util/Multithreading.h:18-48); the baseline way to combine is
// an NCCL all-gather followed by a merge kernel (flatnav_b200/distributed.py keeps that path).  Here the gather is
// not a separate collective: ONE kernel pushes this rank's [Q,K] lists into every peer's gather buffer with plain
// stores through CUDA-IPC-mapped peer pointers (NVLink 5 / NVSwitch), publishes an epoch flag to each peer
// (st.release.sys), waits for the peers' flags (ld.acquire.sys) and runs the k-way merge — transfer, barrier and
// merge in one launch, no NCCL call on the data path.
//
// Buffers (one cudaMalloc per rank, shared through one cudaIpcMemHandle): two halves (epoch parity) of
// [world][max_Q * max_K] distances + labels, then flags[2][16].  A rank can be at most one epoch ahead of a peer
// (it cannot pass the wait of epoch n+1 before the peer has finished merging epoch n and signalled n+1), so two
// halves are enough.
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include "../../include/flatnav_b200.h"
#include "fnb_internal.h"

#define FNB_MAX_RANKS 16

struct fnb_exchange {
  int device = 0, rank = 0, world = 1;
  int64_t max_q = 0;
  int max_k = 0;
  size_t half_elems = 0;  // world * max_q * max_k
  size_t bytes = 0;
  unsigned char* local = nullptr;
  unsigned char* peer[FNB_MAX_RANKS] = {};
  bool attached = false;
  uint32_t epoch = 0;
  unsigned int* ctl = nullptr;  // [0] done counter, [1] error flag
};

namespace fnb {

struct ExchangeParams {
  float* dist[FNB_MAX_RANKS];        // gather buffer (this epoch's half) of every rank; [rank] is the local one
  int32_t* label[FNB_MAX_RANKS];
  uint32_t* flags[FNB_MAX_RANKS];    // flags[r][s]: rank s has delivered its lists of `epoch` to rank r
  unsigned int* ctl;
  float* out_dist;
  int32_t* out_label;
  uint32_t rank, world, Q, K, epoch;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// The grid must be fully co-resident (blocks that already wait for the peers must not keep the remaining pushes of
// this rank from being scheduled): the launcher sizes it with the occupancy API.
__global__ void __launch_bounds__(256) exchange_merge_kernel(const ExchangeParams p) {
  const size_t n = (size_t)p.Q * p.K;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
  // ---- 1. push: this rank's lists (written by the traversal kernel into slot [rank] of the local buffer) go into
  //         slot [rank] of every peer's buffer, straight over NVLink ----
  const float* src_d = p.dist[p.rank] + (size_t)p.rank * n;
  const int32_t* src_l = p.label[p.rank] + (size_t)p.rank * n;
  for (uint32_t r = 0; r < p.world; r++) {
    if (r == p.rank) continue;
    float* dd = p.dist[r] + (size_t)p.rank * n;
    int32_t* dl = p.label[r] + (size_t)p.rank * n;
    for (size_t i = tid; i < n; i += nthreads) {
      dd[i] = src_d[i];
      dl[i] = src_l[i];
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ unsigned int s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(p.ctl, 1u) == gridDim.x - 1u) ? 1u : 0u;
  __syncthreads();
  // ---- 2. signal (the block that saw every other block's pushes) and wait ----
  if (s_last && threadIdx.x < p.world && threadIdx.x != p.rank) {
    __threadfence_system();
    st_release_sys(p.flags[threadIdx.x] + p.rank, p.epoch);
  }
  __shared__ unsigned int s_timeout;
  if (threadIdx.x == 0) s_timeout = 0u;
  __syncthreads();
  if (threadIdx.x < p.world && threadIdx.x != p.rank) {
    const uint32_t* f = p.flags[p.rank] + threadIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(f) - p.epoch) < 0) {
      if (globaltimer_ns() - t0 > 20000000000ull) {  // 20 s: a peer died; report instead of hanging the GPU
        atomicExch(p.ctl + 1, 1u);
        s_timeout = 1u;
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  const bool stale = s_timeout != 0u;  // some peer's lists never arrived: do not merge what an older epoch left there
  // ---- 3. k-way merge of the `world` sorted lists of each query (one warp per query, ties -> lower label) ----
  const int lane = threadIdx.x & 31;
  const size_t warps = nthreads >> 5;
  for (size_t qi = tid >> 5; qi < p.Q; qi += warps) {
    const float* gd = p.dist[p.rank];
    const int32_t* gl = p.label[p.rank];
    const size_t base = (size_t)lane * n + qi * p.K;
    uint32_t head = 0;
    auto load = [&](uint32_t h) -> uint64_t {
      if ((uint32_t)lane >= p.world || h >= p.K) return ~0ull;
      const int32_t lab = __ldcv(gl + base + h);
      if (lab < 0) return ~0ull;
      return ((uint64_t)ord_f32(__ldcv(gd + base + h)) << 32) | (uint32_t)lab;
    };
    uint64_t cur = load(0);
    for (uint32_t j = 0; j < p.K; j++) {
      uint64_t best = cur;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const uint64_t o = shfl64(best, lane ^ off);
        best = o < best ? o : best;
      }
      if (lane == 0) {
        const bool none = stale || best == ~0ull;
        p.out_dist[qi * p.K + j] = none ? __int_as_float(0x7f800000) : unord_f32((uint32_t)(best >> 32));
        p.out_label[qi * p.K + j] = none ? -1 : (int32_t)(uint32_t)best;
      }
      if (best != ~0ull && cur == best) {
        head++;
        cur = load(head);
      }
    }
  }
}

}  // namespace fnb

using namespace fnb;

#define EX_CU(call)                                                                                        \
  do {                                                                                                     \
    cudaError_t e__ = (call);                                                                              \
    if (e__ != cudaSuccess)                                                                                \
      return fail(FNB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

extern "C" {

int fnb_exchange_create(int device, int rank, int world, int64_t max_q, int max_k, fnb_exchange** out) {
  if (!out) return fail(FNB_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (world < 1 || world > FNB_MAX_RANKS || rank < 0 || rank >= world)
    return fail(FNB_ERR_INVALID_ARG, "bad rank %d / world %d (at most %d ranks)", rank, world, FNB_MAX_RANKS);
  if (max_q <= 0 || max_k <= 0) return fail(FNB_ERR_INVALID_ARG, "max_q and max_k must be positive");
  int prev = 0;
  cudaGetDevice(&prev);
  EX_CU(cudaSetDevice(device));
  fnb_exchange* ex = new fnb_exchange();
  ex->device = device;
  ex->rank = rank;
  ex->world = world;
  ex->max_q = max_q;
  ex->max_k = max_k;
  ex->half_elems = (size_t)world * (size_t)max_q * (size_t)max_k;
  ex->bytes = 2 * ex->half_elems * 8 + 2 * FNB_MAX_RANKS * 4;
  cudaError_t e = cudaMalloc(&ex->local, ex->bytes);
  if (e == cudaSuccess) e = cudaMemset(ex->local, 0, ex->bytes);
  if (e == cudaSuccess) e = cudaMalloc(&ex->ctl, 64);
  if (e == cudaSuccess) e = cudaMemset(ex->ctl, 0, 64);
  cudaSetDevice(prev);
  if (e != cudaSuccess) {
    delete ex;
    return fail(FNB_ERR_CUDA, "exchange buffer allocation failed: %s", cudaGetErrorString(e));
  }
  ex->peer[rank] = ex->local;
  ex->attached = world == 1;
  *out = ex;
  return FNB_OK;
}

int fnb_exchange_handle(fnb_exchange* ex, void* handle_out64) {
  if (!ex || !handle_out64) return fail(FNB_ERR_INVALID_ARG, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == FNB_IPC_HANDLE_BYTES, "handle size");
  int prev = 0;
  cudaGetDevice(&prev);
  EX_CU(cudaSetDevice(ex->device));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ex->local);
  cudaSetDevice(prev);
  if (e != cudaSuccess) return fail(FNB_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  memcpy(handle_out64, &h, sizeof(h));
  return FNB_OK;
}

int fnb_exchange_attach(fnb_exchange* ex, const void* handles) {
  if (!ex || !handles) return fail(FNB_ERR_INVALID_ARG, "NULL argument");
  if (ex->attached) return FNB_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  EX_CU(cudaSetDevice(ex->device));
  for (int r = 0; r < ex->world; r++) {
    if (r == ex->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const unsigned char*)handles + (size_t)r * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaSetDevice(prev);
      return fail(FNB_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
    }
    ex->peer[r] = (unsigned char*)ptr;
  }
  cudaSetDevice(prev);
  ex->attached = true;
  return FNB_OK;
}

void fnb_exchange_free(fnb_exchange* ex) {
  if (!ex) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(ex->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < ex->world; r++)
    if (r != ex->rank && ex->peer[r]) cudaIpcCloseMemHandle(ex->peer[r]);
  cudaFree(ex->local);
  cudaFree(ex->ctl);
  cudaSetDevice(prev);
  delete ex;
}

int fnb_search_sharded(fnb_index* ix, fnb_exchange* ex, const void* d_queries, int64_t Q, int K, int ef_search,
                       int num_initializations, float* d_out_dist, int32_t* d_out_label, void* cuda_stream) {
  if (!ix || !ex) return fail(FNB_ERR_INVALID_ARG, "NULL argument");
  if (!ex->attached) return fail(FNB_ERR_INVALID_ARG, "exchange is not attached to its peers yet");
  if (Q <= 0 || Q > ex->max_q || K <= 0 || K > ex->max_k)
    return fail(FNB_ERR_INVALID_ARG, "Q=%lld K=%d exceed the exchange capacity (%lld, %d)", (long long)Q, K,
                (long long)ex->max_q, ex->max_k);
  if (!d_out_dist || !d_out_label) return fail(FNB_ERR_INVALID_ARG, "NULL buffer");
  if (ix->replicas[0].device != ex->device) return fail(FNB_ERR_INVALID_ARG, "index and exchange live on different devices");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  // The epoch is committed only when both kernels of this call are enqueued: a call that fails before that (bad
  // arguments, a plan that does not fit, a launch error) must not leave this rank one epoch ahead of peers it never
  // signalled.
  const uint32_t epoch = ex->epoch + 1u;
  const size_t half = (epoch & 1u) * ex->half_elems;
  const size_t n = (size_t)Q * K;
  ExchangeParams p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < ex->world; r++) {
    p.dist[r] = reinterpret_cast<float*>(ex->peer[r]) + half;
    p.label[r] = reinterpret_cast<int32_t*>(ex->peer[r] + 2 * ex->half_elems * 4) + half;
    p.flags[r] = reinterpret_cast<uint32_t*>(ex->peer[r] + 2 * ex->half_elems * 8) + (epoch & 1u) * FNB_MAX_RANKS;
  }
  p.ctl = ex->ctl;
  p.out_dist = d_out_dist;
  p.out_label = d_out_label;
  p.rank = (uint32_t)ex->rank;
  p.world = (uint32_t)ex->world;
  p.Q = (uint32_t)Q;
  p.K = (uint32_t)K;
  p.epoch = epoch;
  // traversal of the local shard, results straight into slot [rank] of the local gather buffer
  int rc = fnb_search_device(ix, 0, d_queries, Q, K, ef_search, num_initializations, p.dist[ex->rank] + (size_t)ex->rank * n,
                             p.label[ex->rank] + (size_t)ex->rank * n, nullptr, nullptr, cuda_stream);
  if (rc != FNB_OK) return rc;
  int prev = 0;
  cudaGetDevice(&prev);
  EX_CU(cudaSetDevice(ex->device));
  EX_CU(cudaMemsetAsync(ex->ctl, 0, 4, s));
  int per_sm = 0;
  EX_CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, exchange_merge_kernel, 256, 0));
  const int sms = ix->replicas[0].num_sms;
  int grid = sms * (per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm));  // co-resident by construction
  const long long need = ((long long)Q * 32 + 255) / 256;
  if (grid > need) grid = (int)(need < 1 ? 1 : need);
  exchange_merge_kernel<<<grid, 256, 0, s>>>(p);
  cudaError_t e = cudaGetLastError();
  cudaSetDevice(prev);
  if (e != cudaSuccess) return fail(FNB_ERR_CUDA, "exchange kernel launch failed: %s", cudaGetErrorString(e));
  ex->epoch = epoch;
  return FNB_OK;
}

int fnb_exchange_status(fnb_exchange* ex) {
  if (!ex) return fail(FNB_ERR_INVALID_ARG, "NULL argument");
  unsigned int ctl[2] = {0, 0};
  int prev = 0;
  cudaGetDevice(&prev);
  EX_CU(cudaSetDevice(ex->device));
  cudaError_t e = cudaMemcpy(ctl, ex->ctl, 8, cudaMemcpyDeviceToHost);
  cudaSetDevice(prev);
  if (e != cudaSuccess) return fail(FNB_ERR_CUDA, "cudaMemcpy failed: %s", cudaGetErrorString(e));
  if (ctl[1]) {
    // the flag is consumed by this read; the affected calls returned +inf / -1 everywhere.  The ranks' epochs may have
    // diverged (the peer that failed did not advance): re-create the exchange before searching again.
    int prev2 = 0;
    cudaGetDevice(&prev2);
    cudaSetDevice(ex->device);
    cudaMemset(ex->ctl + 1, 0, 4);
    cudaSetDevice(prev2);
    return fail(FNB_ERR_CUDA, "a peer rank did not deliver its results within 20 s; the outputs of that search are "
                              "+inf / -1 and the exchange must be re-created");
  }
  return FNB_OK;
}

}  // extern "C"
