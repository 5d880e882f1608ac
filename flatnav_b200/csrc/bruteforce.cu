// Exact scan (ground truth / exact re-rank): top-K by (distance, node id) over every live node, with the
// same distance arithmetic (lane order, fused multiply-add) as the traversal kernel, so the values are
// bit-identical to what fnb_search returns for the same (query, node) pair.
//
// The reference has no brute force of its own (recall in its tests is computed against externally
// supplied ground truth, python-bindings/unit_tests/test_utils.py:57-91); semantics are pinned by
// oracle/flatnav_oracle.cpp `ora_bruteforce` and a numpy float64 scan in tests/.
//
// CUDA-core version: one warp per query, a CTA of 8 warps streams the database through a shared-memory
// tile that all 8 queries reuse.
#include "../../include/flatnav_b200.h"
#include <cstdlib>
#include <cstring>

#include "bf_common.cuh"

namespace fnb {

template <int DT, int METRIC, int G, int CH>
__global__ void __launch_bounds__(BF_WARPS * 32) bruteforce_kernel(const BfParams p) {
  typedef Arith<DT, METRIC> A;
  extern __shared__ __align__(16) unsigned char bf_smem[];
  uint4* tile = reinterpret_cast<uint4*>(bf_smem);                                  // [BF_TILE_ROWS][nchunks]
  uint64_t* lists = reinterpret_cast<uint64_t*>(tile + (size_t)BF_TILE_ROWS * p.nchunks);  // [BF_WARPS][Kcap]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / G, pos = lane % G;
  constexpr int RPI = 32 / G;
  uint64_t* list = lists + (size_t)warp * p.Kcap;
  const uint32_t slot = blockIdx.x * BF_WARPS + warp;
  const bool active = slot < p.Q;
  const uint32_t qi = (active && p.qmap) ? p.qmap[slot] : slot;

  SearchParams sp;  // only the fields load_query_chunk reads
  sp.queries = p.queries;
  sp.dim = p.dim;
  sp.nchunks = p.nchunks;
  sp.query_vec_ok = p.query_vec_ok;
  sp.query_pitch_chunks = 0;
  uint4 q[CH];
#pragma unroll
  for (int k = 0; k < CH; k++) q[k] = active ? load_query_chunk<DT>(sp, qi, (uint32_t)(k * G + pos)) : make_uint4(0, 0, 0, 0);

  uint32_t len = 0;
  for (uint32_t base = 0; base < p.N; base += BF_TILE_ROWS) {
    const uint32_t rows = min((uint32_t)BF_TILE_ROWS, p.N - base);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < rows * p.nchunks; i += blockDim.x) {
      const uint32_t r = i / p.nchunks, c = i - r * p.nchunks;
      tile[i] = ldg_stream(p.vec + (size_t)(base + r) * p.stride + c);
    }
    __syncthreads();
    if (!active) continue;
    for (uint32_t r0 = 0; r0 < rows; r0 += RPI) {
      const uint32_t r = r0 + g;
      const bool ok = r < rows;
      typename A::acc_t acc = 0;
#pragma unroll
      for (int k = 0; k < CH; k++) {
        const uint32_t chunk = (uint32_t)(k * G + pos);
        const uint4 x = (ok && chunk < p.nchunks) ? tile[(size_t)r * p.nchunks + chunk] : make_uint4(0, 0, 0, 0);
        A::step(acc, q[k], x);
      }
#pragma unroll
      for (int off = G / 2; off > 0; off >>= 1) acc = A::combine(acc, shfl_xor_t(acc, off));
      const uint64_t key = ((uint64_t)ord_f32(A::finish(acc)) << 32) | (uint64_t)(base + r);
      const uint64_t worst = len ? list[len - 1] : 0ull;
      const bool cand = ok && pos == 0 && (len < p.K || key < worst);
      unsigned cm = __ballot_sync(FNB_FULL, cand);
      for (; cm; cm &= cm - 1) {  // rare after the first few tiles: sequential warp-parallel insertion
        warp_topk_insert(list, len, p.K, shfl64(key, __ffs(cm) - 1), lane);
      }
    }
  }
  if (!active) return;
  __syncwarp();
  for (uint32_t i = lane; i < p.K; i += 32) {
    float od = __int_as_float(0x7f800000);
    int32_t ol = -1;
    if (i < len) {
      const uint64_t e = list[i];
      od = unord_f32((uint32_t)(e >> 32));
      ol = __ldg(p.labels + (uint32_t)e);
    }
    p.out_dist[(size_t)qi * p.K + i] = od;
    p.out_label[(size_t)qi * p.K + i] = ol;
  }
}

template <int DT, int METRIC, int G, int CH>
static cudaError_t launch_bf(const BfParams& p, cudaStream_t s) {
  auto kern = bruteforce_kernel<DT, METRIC, G, CH>;
  const size_t smem = (size_t)BF_TILE_ROWS * p.nchunks * 16 + (size_t)BF_WARPS * p.Kcap * 8;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<(p.Q + BF_WARPS - 1) / BF_WARPS, BF_WARPS * 32, smem, s>>>(p);
  return cudaGetLastError();
}

template <int DT, int METRIC>
static cudaError_t bf_gc(const fnb_index* ix, const BfParams& p, cudaStream_t s) {
  const int ch = fnb_chunks_per_lane(ix->nchunks, ix->G);
  if (ix->G == 4) {
    if (ch <= 1) return launch_bf<DT, METRIC, 4, 1>(p, s);
    return launch_bf<DT, METRIC, 4, 2>(p, s);
  }
  if (ix->G == 8) {
    switch (ch) {
      case 1: return launch_bf<DT, METRIC, 8, 1>(p, s);
      case 2: return launch_bf<DT, METRIC, 8, 2>(p, s);
      case 3: return launch_bf<DT, METRIC, 8, 3>(p, s);
      default: return launch_bf<DT, METRIC, 8, 4>(p, s);
    }
  }
  if (ch <= 2) return launch_bf<DT, METRIC, 32, 2>(p, s);
  if (ch <= 4) return launch_bf<DT, METRIC, 32, 4>(p, s);
  if (ch <= 8) return launch_bf<DT, METRIC, 32, 8>(p, s);
  return launch_bf<DT, METRIC, 32, 16>(p, s);
}

cudaError_t launch_exact_scan(const fnb_index* ix, const BfParams& p, cudaStream_t s) {
  const bool ip = ix->h.metric == FNB_METRIC_IP;
  switch (ix->h.data_type) {
    case FNB_DTYPE_FLOAT32: return ip ? bf_gc<DT_F32, M_IP>(ix, p, s) : bf_gc<DT_F32, M_L2>(ix, p, s);
    case FNB_DTYPE_UINT8: return ip ? bf_gc<DT_U8, M_IP>(ix, p, s) : bf_gc<DT_U8, M_L2>(ix, p, s);
    default: return ip ? bf_gc<DT_I8, M_IP>(ix, p, s) : bf_gc<DT_I8, M_L2>(ix, p, s);
  }
}

static thread_local BfRun g_last_bf;

}  // namespace fnb

using namespace fnb;

extern "C" int fnb_bruteforce(fnb_index* ix, const void* queries, int64_t Q, int K, float* out_dist,
                              int32_t* out_label) {
  if (!ix) return fail(FNB_ERR_INVALID_ARG, "index is NULL");
  if (K <= 0 || Q < 0) return fail(FNB_ERR_INVALID_ARG, "bad K or Q");
  if (Q == 0) return FNB_OK;
  if (!queries || !out_dist || !out_label) return fail(FNB_ERR_INVALID_ARG, "NULL buffer");
  if (K > 2048) return fail(FNB_ERR_UNSUPPORTED, "brute force supports K <= 2048");
  if (Q >= (1ll << 31)) return fail(FNB_ERR_UNSUPPORTED, "more than 2^31 queries in one call");
  fnb::ExclusiveLock lock(ix->mu);
  fnb::quiesce(ix);
  Replica& r = ix->replicas[0];
  const Header& h = ix->h;
  int prev = 0;
  cudaGetDevice(&prev);
#define BF_CU(call)                                                                                   \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess) {                                                                         \
      cudaSetDevice(prev);                                                                            \
      return fail(FNB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));                     \
    }                                                                                                 \
  } while (0)
  BF_CU(cudaSetDevice(r.device));
  struct DevBufs {  // freed on every way out, error returns included
    unsigned char *q = nullptr, *d = nullptr, *l = nullptr;
    ~DevBufs() {
      cudaFree(q);
      cudaFree(d);
      cudaFree(l);
    }
  } bufs;
  const size_t qb = (size_t)Q * h.data_size, ob = (size_t)Q * K * 4;
  BF_CU(cudaMalloc(&bufs.q, qb));
  BF_CU(cudaMalloc(&bufs.d, ob));
  BF_CU(cudaMalloc(&bufs.l, ob));
  unsigned char *d_q = bufs.q, *d_d = bufs.d, *d_l = bufs.l;
  BF_CU(cudaMemcpyAsync(d_q, queries, qb, cudaMemcpyHostToDevice, r.stream));
  // FNB_BF_MODE=exact|tensor forces a path (tests, profiling); default: the tcgen05 filter + exact re-rank
  // whenever the problem is large enough to fill tensor-core tiles.
  const char* mode = getenv("FNB_BF_MODE");
  const bool force_exact = mode && !strcmp(mode, "exact");
  const bool force_tensor = mode && !strcmp(mode, "tensor");
  BfRun run;
  int rc = FNB_OK;
  if (!force_exact && tensor_path_supported(ix, Q, K) &&
      (force_tensor || ((double)Q * (double)h.cur_nodes >= 1e8 && h.cur_nodes >= 4096))) {
    run.path = 1;
    rc = bruteforce_tensor(ix, r, d_q, Q, K, reinterpret_cast<float*>(d_d), reinterpret_cast<int32_t*>(d_l), &run);
  } else if (force_tensor) {
    rc = fail(FNB_ERR_UNSUPPORTED, "FNB_BF_MODE=tensor but the tensor path does not support this problem (K=%d)", K);
  } else {
    BfParams p;
    memset(&p, 0, sizeof(p));
    p.vec = r.vec;
    p.labels = r.labels;
    p.queries = d_q;
    p.qmap = nullptr;
    p.out_dist = reinterpret_cast<float*>(d_d);
    p.out_label = reinterpret_cast<int32_t*>(d_l);
    p.N = (uint32_t)h.cur_nodes;
    p.dim = (uint32_t)h.dim;
    p.nchunks = ix->nchunks;
    p.stride = ix->stride;
    p.Q = (uint32_t)Q;
    p.K = (uint32_t)K;
    p.Kcap = ((uint32_t)K + 31u) & ~31u;
    p.query_vec_ok = (h.data_size % FNB_CHUNK_BYTES) == 0 ? 1u : 0u;
    cudaEvent_t e0, e1;
    BF_CU(cudaEventCreate(&e0));
    BF_CU(cudaEventCreate(&e1));
    BF_CU(cudaEventRecord(e0, r.stream));
    BF_CU(launch_exact_scan(ix, p, r.stream));
    BF_CU(cudaEventRecord(e1, r.stream));
    BF_CU(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&run.rescan_ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  if (rc == FNB_OK) {
    BF_CU(cudaMemcpyAsync(out_dist, d_d, ob, cudaMemcpyDeviceToHost, r.stream));
    BF_CU(cudaMemcpyAsync(out_label, d_l, ob, cudaMemcpyDeviceToHost, r.stream));
    BF_CU(cudaStreamSynchronize(r.stream));
  }
  cudaSetDevice(prev);
  g_last_bf = run;
  return rc;
}

extern "C" int fnb_bruteforce_stats(fnb_bf_stats* out) {
  if (!out) return fail(FNB_ERR_INVALID_ARG, "NULL argument");
  memset(out, 0, sizeof(*out));
  out->path = g_last_bf.path;
  out->n_unsafe = g_last_bf.n_unsafe;
  out->n_candidates = g_last_bf.n_candidates;
  out->prep_ms = g_last_bf.prep_ms;
  out->gemm_ms = g_last_bf.gemm_ms;
  out->rerank_ms = g_last_bf.rerank_ms;
  out->rescan_ms = g_last_bf.rescan_ms;
  out->gemm_flops = g_last_bf.gemm_flops;
  return FNB_OK;
}
