// Exact re-rank of caller-supplied candidate lists ("benchmark driver parity ... and exact
// re-rank option").  The reference has no re-rank of its own — its benchmark driver (experiments/run-benchmark.py:
// 38-124) takes whatever `search_single` returns — so the semantics are this engine's: for every query, evaluate the
// distance to each candidate with the traversal kernel's arithmetic (bit-identical to what fnb_search / fnb_bruteforce
// report for the same (query, node) pair), drop unknown and repeated candidates, and return the K best by
// (distance, node id) with the nodes' label fields.  Candidates are node labels (what search() hands out) or node ids.
//
// Labels -> nodes: labels are arbitrary int32 values stored with the nodes (Index.h:262-272), so a sorted
// (label, node) table is built on first use and kept until the index changes; when every label equals its node id
// (the default before a re-ordering) no table is needed.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/flatnav_b200.h"
#include "bf_common.cuh"

struct fnb_label_map {
  bool identity = false;
  int32_t* sorted_label = nullptr;  // device, ascending
  uint32_t* sorted_node = nullptr;  // device
  uint64_t n = 0;
};

void fnb_label_map_free(fnb_label_map* m) {
  if (!m) return;
  cudaFree(m->sorted_label);
  cudaFree(m->sorted_node);
  delete m;
}

namespace fnb {

struct RerankCallParams {
  SearchParams sp;  // vec / stride / nchunks / dim / queries for the distance code
  const int32_t* __restrict__ labels;
  const int32_t* __restrict__ cand;          // [Q][C]
  const int32_t* __restrict__ sorted_label;  // null: candidates are node ids (or labels == node ids)
  const uint32_t* __restrict__ sorted_node;
  float* __restrict__ out_dist;
  int32_t* __restrict__ out_label;
  uint32_t n_sorted, Q, C, K, Kcap;
};

// one warp per query, 4 warps per CTA
template <int DT, int METRIC, int G, int CH, bool EXACT>
__global__ void __launch_bounds__(128) rerank_kernel(const RerankCallParams p) {
  extern __shared__ __align__(16) unsigned char rr_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t* list = reinterpret_cast<uint64_t*>(rr_smem) + (size_t)warp * (p.Kcap + 16);
  uint32_t* s_ids = reinterpret_cast<uint32_t*>(list + p.Kcap);
  const uint32_t qi = blockIdx.x * 4 + warp;
  if (qi >= p.Q) return;
  const int pos = lane % G;
  uint4 q[CH];
#pragma unroll
  for (int k = 0; k < CH; k++) q[k] = load_query_chunk<DT>(p.sp, qi, (uint32_t)(k * G + pos));
  uint32_t len = 0;
  for (uint32_t c0 = 0; c0 < p.C; c0 += 32) {
    const uint32_t c = c0 + lane;
    int32_t v = c < p.C ? p.cand[(size_t)qi * p.C + c] : -1;
    uint32_t node = 0xffffffffu;
    if (c < p.C) {
      if (p.sorted_label) {  // lower_bound of the label, then an exact match
        uint32_t lo = 0, hi = p.n_sorted;
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (__ldg(p.sorted_label + mid) < v) lo = mid + 1;
          else hi = mid;
        }
        if (lo < p.n_sorted && __ldg(p.sorted_label + lo) == v) node = __ldg(p.sorted_node + lo);
      } else if (v >= 0) {
        node = (uint32_t)v;
      }
    }
    const bool valid = node < p.sp.N;
    const float d = batch_distance<DT, METRIC, G, CH, EXACT>(p.sp, q, valid ? node : 0u, valid, s_ids, lane, false);
    const uint64_t key = ((uint64_t)ord_f32(d) << 32) | (uint64_t)node;
    for (unsigned cm = __ballot_sync(FNB_FULL, valid); cm; cm &= cm - 1) {
      const uint64_t kx = shfl64(key, __ffs(cm) - 1);
      // a candidate listed twice has the same (distance, node) key: skip it
      bool dup = false;
      for (uint32_t i = lane; i < len; i += 32) dup |= list[i] == kx;
      if (__any_sync(FNB_FULL, dup)) continue;
      warp_topk_insert(list, len, p.K, kx, lane);
    }
  }
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
static cudaError_t launch_rr(const RerankCallParams& p, cudaStream_t s) {
  const bool exact = p.sp.nchunks == (uint32_t)(G * CH);
  auto kern = exact ? rerank_kernel<DT, METRIC, G, CH, true> : rerank_kernel<DT, METRIC, G, CH, false>;
  const size_t smem = (size_t)4 * (p.Kcap + 16) * 8;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<(p.Q + 3) / 4, 128, smem, s>>>(p);
  return cudaGetLastError();
}

template <int DT, int METRIC>
static cudaError_t rr_gc(const fnb_index* ix, const RerankCallParams& p, cudaStream_t s) {
  const int ch = fnb_chunks_per_lane(ix->nchunks, ix->G);
  if (ix->G == 4) {
    if (ch <= 1) return launch_rr<DT, METRIC, 4, 1>(p, s);
    return launch_rr<DT, METRIC, 4, 2>(p, s);
  }
  if (ix->G == 8) {
    switch (ch) {
      case 1: return launch_rr<DT, METRIC, 8, 1>(p, s);
      case 2: return launch_rr<DT, METRIC, 8, 2>(p, s);
      case 3: return launch_rr<DT, METRIC, 8, 3>(p, s);
      default: return launch_rr<DT, METRIC, 8, 4>(p, s);
    }
  }
  if (ch <= 2) return launch_rr<DT, METRIC, 32, 2>(p, s);
  if (ch <= 4) return launch_rr<DT, METRIC, 32, 4>(p, s);
  if (ch <= 8) return launch_rr<DT, METRIC, 32, 8>(p, s);
  return launch_rr<DT, METRIC, 32, 16>(p, s);
}

static cudaError_t launch_rerank(const fnb_index* ix, const RerankCallParams& p, cudaStream_t s) {
  const bool ip = ix->h.metric == FNB_METRIC_IP;
  switch (ix->h.data_type) {
    case FNB_DTYPE_FLOAT32: return ip ? rr_gc<DT_F32, M_IP>(ix, p, s) : rr_gc<DT_F32, M_L2>(ix, p, s);
    case FNB_DTYPE_UINT8: return ip ? rr_gc<DT_U8, M_IP>(ix, p, s) : rr_gc<DT_U8, M_L2>(ix, p, s);
    default: return ip ? rr_gc<DT_I8, M_IP>(ix, p, s) : rr_gc<DT_I8, M_L2>(ix, p, s);
  }
}

}  // namespace fnb

using namespace fnb;

#define RR_CU(call)                                                                                           \
  do {                                                                                                        \
    cudaError_t e__ = (call);                                                                                 \
    if (e__ != cudaSuccess)                                                                                   \
      return fail(FNB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

namespace {
struct DevPtr {
  void* p = nullptr;
  ~DevPtr() { cudaFree(p); }
};
struct DevScope {
  int prev = 0;
  DevScope() { cudaGetDevice(&prev); }
  ~DevScope() { cudaSetDevice(prev); }
};
struct StreamScope {
  cudaStream_t s = nullptr;
  ~StreamScope() {
    if (s) cudaStreamDestroy(s);
  }
};
}  // namespace

// caller holds the index lock (shared is enough: the map has its own mutex) and has selected the replica's device
static int ensure_label_map(fnb_index* ix, const fnb_label_map** out) {
  std::lock_guard<std::mutex> lk(ix->aux_mu);
  if (ix->label_map && ix->label_map->n == ix->h.cur_nodes) {
    *out = ix->label_map;
    return FNB_OK;
  }
  fnb_label_map_free(ix->label_map);
  ix->label_map = nullptr;
  const Replica& r = ix->replicas[0];
  const uint64_t n = ix->h.cur_nodes;
  std::vector<int32_t> lab(n);
  if (n) RR_CU(cudaMemcpy(lab.data(), r.labels, n * 4, cudaMemcpyDeviceToHost));
  fnb_label_map* m = new fnb_label_map();
  m->n = n;
  m->identity = true;
  for (uint64_t i = 0; i < n; i++)
    if (lab[i] != (int32_t)i) {
      m->identity = false;
      break;
    }
  if (!m->identity) {
    std::vector<uint32_t> node(n);
    std::iota(node.begin(), node.end(), 0u);
    // equal labels (the reference does not forbid them): the lowest node id answers for the label
    std::sort(node.begin(), node.end(), [&](uint32_t a, uint32_t b) { return lab[a] != lab[b] ? lab[a] < lab[b] : a < b; });
    std::vector<int32_t> sl(n);
    for (uint64_t i = 0; i < n; i++) sl[i] = lab[node[i]];
    cudaError_t e = cudaMalloc((void**)&m->sorted_label, n * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->sorted_node, n * 4);
    if (e == cudaSuccess) e = cudaMemcpy(m->sorted_label, sl.data(), n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(m->sorted_node, node.data(), n * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      fnb_label_map_free(m);
      return fail(FNB_ERR_CUDA, "cannot build the label table: %s", cudaGetErrorString(e));
    }
  }
  ix->label_map = m;
  *out = m;
  return FNB_OK;
}

extern "C" int fnb_rerank(fnb_index* ix, const void* queries, int64_t Q, const int32_t* candidates, int C,
                          int candidates_are_labels, int K, float* out_dist, int32_t* out_label) {
  if (!ix) return fail(FNB_ERR_INVALID_ARG, "index is NULL");
  if (K <= 0 || Q < 0 || C <= 0) return fail(FNB_ERR_INVALID_ARG, "bad K, Q or candidate count");
  if (Q == 0) return FNB_OK;
  if (!queries || !candidates || !out_dist || !out_label) return fail(FNB_ERR_INVALID_ARG, "NULL buffer");
  if (K > 2048) return fail(FNB_ERR_UNSUPPORTED, "re-rank supports K <= 2048");
  if (Q >= (1ll << 31)) return fail(FNB_ERR_UNSUPPORTED, "more than 2^31 queries in one call");
  SharedLock lock(ix->mu);
  const Header& h = ix->h;
  const Replica& r = ix->replicas[0];
  DevScope scope;
  RR_CU(cudaSetDevice(r.device));
  const fnb_label_map* map = nullptr;
  if (candidates_are_labels) {
    const int rc = ensure_label_map(ix, &map);
    if (rc != FNB_OK) return rc;
  }
  StreamScope st;
  RR_CU(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
  DevPtr d_q, d_c, d_d, d_l;
  const size_t qb = (size_t)Q * h.data_size, cb = (size_t)Q * C * 4, ob = (size_t)Q * K * 4;
  RR_CU(cudaMalloc(&d_q.p, qb));
  RR_CU(cudaMalloc(&d_c.p, cb));
  RR_CU(cudaMalloc(&d_d.p, ob));
  RR_CU(cudaMalloc(&d_l.p, ob));
  RR_CU(cudaMemcpyAsync(d_q.p, queries, qb, cudaMemcpyHostToDevice, st.s));
  RR_CU(cudaMemcpyAsync(d_c.p, candidates, cb, cudaMemcpyHostToDevice, st.s));
  RerankCallParams p;
  memset(&p, 0, sizeof(p));
  p.sp.vec = r.vec;
  p.sp.queries = d_q.p;
  p.sp.N = (uint32_t)h.cur_nodes;
  p.sp.dim = (uint32_t)h.dim;
  p.sp.nchunks = ix->nchunks;
  p.sp.stride = ix->stride;
  p.sp.query_vec_ok = (h.data_size % FNB_CHUNK_BYTES) == 0 ? 1u : 0u;
  p.sp.lines_per_row = (ix->stride * FNB_CHUNK_BYTES + 127u) / 128u;
  p.labels = r.labels;
  p.cand = static_cast<const int32_t*>(d_c.p);
  if (map && !map->identity) {
    p.sorted_label = map->sorted_label;
    p.sorted_node = map->sorted_node;
    p.n_sorted = (uint32_t)map->n;
  }
  p.out_dist = static_cast<float*>(d_d.p);
  p.out_label = static_cast<int32_t*>(d_l.p);
  p.Q = (uint32_t)Q;
  p.C = (uint32_t)C;
  p.K = (uint32_t)K;
  p.Kcap = ((uint32_t)K + 31u) & ~31u;
  RR_CU(launch_rerank(ix, p, st.s));
  RR_CU(cudaMemcpyAsync(out_dist, d_d.p, ob, cudaMemcpyDeviceToHost, st.s));
  RR_CU(cudaMemcpyAsync(out_label, d_l.p, ob, cudaMemcpyDeviceToHost, st.s));
  RR_CU(cudaStreamSynchronize(st.s));
  return FNB_OK;
}
