// Graph construction on the GPU: batched insertion with the reference's own rules.
//
// Replaces Index::add / addBatch (include/flatnav/index/Index.h:301-378), selectNeighbors (:714-763) and
// connectNeighbors (:765-834).  The reference inserts one node at a time (or T threads under per-node mutexes); here
// a BATCH of nodes is inserted per step:
//   1. search   the traversal kernel (search_kernel.cuh) finds, for every node of the batch, its ef_construction
//               nearest nodes among the nodes inserted by earlier batches (the node's own row in HBM is the query);
//   2. select   build_select_kernel, one warp per new node: the HNSW heuristic of selectNeighbors with
//               M/2 slots (Index.h:373-375: selection_M = max(M/2, 1)): candidates in ascending distance, keep c iff
//               no already-kept s has dist(s, c) < dist(new, c); fewer than M/2 candidates => keep all (:715-717);
//               then writes the new node's links and the back-links: a free slot of the neighbour's row (the
//               reference replaces the first self-loop, :785-793; rows stay packed, so the slot is an atomicAdd on a
//               per-node degree counter), or, if the row is full, an entry in the neighbour's overflow list;
//   3. prune    build_prune_kernel, one warp per node with overflow: old links + newcomers, distances to the node,
//               sort, the same heuristic with M slots, rewrite the row (:794-826; the reference does this once per
//               newcomer, a batch does it once per node).
// Nodes of one batch cannot link to each other (they are not in the graph while it is searched); batches are kept
// <= 1/16 .. 1/8 of the nodes already inserted, so the loss is small: recall of the built graphs is within noise of
// reference-built ones (tests/test_gpu_build.py), and the files are read by the unmodified reference.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/flatnav_b200.h"
#include "fnb_internal.h"

namespace fnb {

#define BUILD_WARPS 4
#define BUILD_TMAX 128  // candidates one prune handles: M old links + up to TMAX - M newcomers

struct BuildParams {
  SearchParams sp;  // vec / stride / nchunks / dim for the distance code; queries = vec, pitch = stride
  uint32_t* adj;
  uint32_t* deg;
  const int32_t* cand_id;  // [b][Kc] node ids ascending by distance, -1 = none
  const float* cand_dist;  // [b][Kc]
  uint32_t* ovf_head;      // [max_nodes] 1 + index of the node's first overflow entry, 0 = none
  uint32_t* ovf_next;      // [cap]
  uint32_t* ovf_src;       // [cap]
  uint32_t* dirty;         // nodes that got their first overflow entry in this batch
  unsigned int* counters;  // [0] overflow entries, [1] dirty nodes, [2] dropped newcomers, [3] prune work counter
  uint32_t M, Msel, Kc, first, b, ovf_cap;
};

template <int DT, int G, int CH>
__device__ __forceinline__ void load_row_as_query(const SearchParams& sp, uint32_t node, int lane, uint4 (&q)[CH]) {
  const int pos = lane % G;
#pragma unroll
  for (int k = 0; k < CH; k++) {
    const uint32_t chunk = (uint32_t)(k * G + pos);
    q[k] = chunk < sp.nchunks ? __ldg(sp.vec + (size_t)node * sp.stride + chunk) : make_uint4(0, 0, 0, 0);
  }
}

// HNSW heuristic over `n` candidates sorted ascending by distance to the centre node (ids in c_id, distances in
// c_dist, shared memory): fills sel[0..limit) and returns how many were kept.
template <int DT, int METRIC, int G, int CH, bool EXACT>
__device__ __forceinline__ uint32_t heuristic_select(const SearchParams& sp, const uint32_t* c_id, const float* c_dist,
                                                     uint32_t n, uint32_t limit, uint32_t* sel, uint32_t* s_ids, int lane) {
  uint32_t ns = 0;
  if (n < limit) {  // Index.h:715-717: nothing to prune
    for (uint32_t i = lane; i < n; i += 32) sel[i] = c_id[i];
    __syncwarp();
    return n;
  }
  for (uint32_t c = 0; c < n && ns < limit; c++) {
    const uint32_t cid = c_id[c];
    const float dqc = c_dist[c];
    bool reject = false;
    if (ns > 0) {
      uint4 q[CH];
      load_row_as_query<DT, G, CH>(sp, cid, lane, q);
      for (uint32_t base = 0; base < ns && !reject; base += 32) {
        const bool valid = base + lane < ns;
        const uint32_t sid = valid ? sel[base + lane] : 0u;
        const float d = batch_distance<DT, METRIC, G, CH, EXACT>(sp, q, sid, valid, s_ids, lane, false);
        reject = __any_sync(FNB_FULL, valid && d < dqc);  // Index.h:742-746
      }
    }
    if (!reject) {
      if (lane == 0) sel[ns] = cid;
      ns++;
      __syncwarp();
    }
  }
  return ns;
}

template <int DT, int METRIC, int G, int CH, bool EXACT>
__global__ void __launch_bounds__(BUILD_WARPS * 32) build_select_kernel(const BuildParams p) {
  extern __shared__ __align__(16) unsigned char bs_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // per warp: c_id[Kc] | c_dist[Kc] | sel[Msel] | s_ids[32]
  uint32_t* c_id = reinterpret_cast<uint32_t*>(bs_smem) + (size_t)warp * (2 * p.Kc + p.Msel + 32);
  float* c_dist = reinterpret_cast<float*>(c_id + p.Kc);
  uint32_t* sel = c_id + 2 * p.Kc;
  uint32_t* s_ids = sel + p.Msel;
  const uint32_t w = blockIdx.x * BUILD_WARPS + warp;
  if (w >= p.b) return;
  const uint32_t u = p.first + w;
  uint32_t n = 0;
  for (uint32_t i = lane; i < p.Kc; i += 32) {
    const int32_t id = p.cand_id[(size_t)w * p.Kc + i];
    c_id[i] = (uint32_t)id;
    c_dist[i] = p.cand_dist[(size_t)w * p.Kc + i];
    n += id >= 0 ? 1u : 0u;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) n += __shfl_xor_sync(FNB_FULL, n, off);
  __syncwarp();
  const uint32_t ns = heuristic_select<DT, METRIC, G, CH, EXACT>(p.sp, c_id, c_dist, n, p.Msel, sel, s_ids, lane);
  // links of the new node (its row was initialised to self-loops), then the back-links
  for (uint32_t i = lane; i < ns; i += 32) {
    const uint32_t s = sel[i];
    p.adj[(size_t)u * p.M + i] = s;
    const uint32_t slot = atomicAdd(p.deg + s, 1u);
    if (slot < p.M) {
      p.adj[(size_t)s * p.M + slot] = u;
    } else {
      const uint32_t e = atomicAdd(p.counters + 0, 1u);
      if (e < p.ovf_cap) {
        p.ovf_src[e] = u;
        const uint32_t prev = atomicExch(p.ovf_head + s, e + 1u);
        p.ovf_next[e] = prev;
        if (prev == 0u) p.dirty[atomicAdd(p.counters + 1, 1u)] = s;
      } else {
        atomicAdd(p.counters + 2, 1u);
      }
    }
  }
  if (lane == 0) p.deg[u] = ns;
}

template <int DT, int METRIC, int G, int CH, bool EXACT>
__global__ void __launch_bounds__(BUILD_WARPS * 32) build_prune_kernel(const BuildParams p) {
  extern __shared__ __align__(16) unsigned char bp_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // per warp: keys[TMAX] u64 | c_id[TMAX] | c_dist[TMAX] | sel[M] | s_ids[32]
  unsigned char* wb = bp_smem + (size_t)warp * (BUILD_TMAX * 16 + (p.M + 32) * 4);
  uint64_t* keys = reinterpret_cast<uint64_t*>(wb);
  uint32_t* c_id = reinterpret_cast<uint32_t*>(keys + BUILD_TMAX);
  float* c_dist = reinterpret_cast<float*>(c_id + BUILD_TMAX);
  uint32_t* sel = c_id + 2 * BUILD_TMAX;
  uint32_t* s_ids = sel + p.M;
  const uint32_t n_dirty = p.counters[1];
  for (;;) {
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(p.counters + 3, 1u);
    w = __shfl_sync(FNB_FULL, w, 0);
    if (w >= n_dirty) break;
    const uint32_t s = p.dirty[w];
    // gather: the M old links, then the newcomers of this batch
    uint32_t n = 0;
    for (uint32_t i = lane; i < p.M; i += 32) c_id[i] = p.adj[(size_t)s * p.M + i];
    n = p.M;
    uint32_t e = p.ovf_head[s];
    uint32_t dropped = 0;
    while (e != 0u) {  // walked by all lanes in step (broadcast loads)
      if (n < BUILD_TMAX) {
        if (lane == 0) c_id[n] = p.ovf_src[e - 1u];
        n++;
      } else {
        dropped++;
      }
      e = p.ovf_next[e - 1u];
    }
    __syncwarp();
    if (lane == 0) {
      p.ovf_head[s] = 0u;
      if (dropped) atomicAdd(p.counters + 2, dropped);
    }
    // distances to s
    uint4 q[CH];
    load_row_as_query<DT, G, CH>(p.sp, s, lane, q);
    for (uint32_t base = 0; base < n; base += 32) {
      const bool valid = base + lane < n;
      const uint32_t id = valid ? c_id[base + lane] : 0u;
      const float d = batch_distance<DT, METRIC, G, CH, EXACT>(p.sp, q, id, valid, s_ids, lane, false);
      if (valid) keys[base + lane] = ((uint64_t)ord_f32(d) << 32) | id;
    }
    __syncwarp();
    // rank sort (n <= 128): position = number of smaller keys; keys are distinct (ids are)
    uint64_t mine[BUILD_TMAX / 32];
    uint32_t rank[BUILD_TMAX / 32];
#pragma unroll
    for (int j = 0; j < BUILD_TMAX / 32; j++) {
      mine[j] = (uint32_t)(j * 32 + lane) < n ? keys[j * 32 + lane] : ~0ull;
      rank[j] = 0;
    }
    for (uint32_t i = 0; i < n; i++) {
      const uint64_t k = keys[i];
#pragma unroll
      for (int j = 0; j < BUILD_TMAX / 32; j++) rank[j] += k < mine[j] ? 1u : 0u;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < BUILD_TMAX / 32; j++) {
      if ((uint32_t)(j * 32 + lane) < n) {
        c_id[rank[j]] = (uint32_t)mine[j];
        c_dist[rank[j]] = unord_f32((uint32_t)(mine[j] >> 32));
      }
    }
    __syncwarp();
    const uint32_t ns = heuristic_select<DT, METRIC, G, CH, EXACT>(p.sp, c_id, c_dist, n, p.M, sel, s_ids, lane);
    for (uint32_t i = lane; i < p.M; i += 32) p.adj[(size_t)s * p.M + i] = i < ns ? sel[i] : s;  // self-loops fill the row
    if (lane == 0) p.deg[s] = ns;
    __syncwarp();
  }
}

// rows of freshly appended nodes: every slot a self-loop (Index.h:262-272), degree 0
__global__ void build_init_rows_kernel(uint32_t* adj, uint32_t* deg, uint32_t first, uint32_t count, uint32_t M) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)count * M) adj[(size_t)first * M + i] = first + (uint32_t)(i / M);
  if (deg && i < count) deg[first + i] = 0;
}
// degree of loaded rows = number of links that are not self-loops (rows are packed: used slots first)
__global__ void build_count_degree_kernel(const uint32_t* adj, uint32_t* deg, uint32_t n, uint32_t M, unsigned int* unpacked) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  uint32_t d = 0;
  bool gap = false;
  for (uint32_t j = 0; j < M; j++) {
    const bool used = adj[(size_t)v * M + j] != v;
    if (used && d != j) gap = true;
    d += used ? 1u : 0u;
  }
  deg[v] = d;
  if (gap) atomicAdd(unpacked, 1u);
}
// expand [n][dim] dense host-layout vectors into padded rows
__global__ void build_pad_rows_kernel(const unsigned char* __restrict__ src, uint32_t data_size, unsigned char* __restrict__ dst,
                                      uint32_t stride_bytes, uint64_t n) {
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (uint64_t r = (((uint64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; r < n; r += warps)
    for (uint32_t b = lane; b < stride_bytes; b += 32) dst[r * stride_bytes + b] = b < data_size ? src[r * data_size + b] : (unsigned char)0;
}

template <int DT, int METRIC, int G, int CH>
static cudaError_t launch_build(const BuildParams& p, int num_sms, cudaStream_t s) {
  const bool exact = p.sp.nchunks == (uint32_t)(G * CH);
  {
    auto kern = exact ? build_select_kernel<DT, METRIC, G, CH, true> : build_select_kernel<DT, METRIC, G, CH, false>;
    const size_t smem = (size_t)BUILD_WARPS * (2 * p.Kc + p.Msel + 32) * 4;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(p.b + BUILD_WARPS - 1) / BUILD_WARPS, BUILD_WARPS * 32, smem, s>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  {
    auto kern = exact ? build_prune_kernel<DT, METRIC, G, CH, true> : build_prune_kernel<DT, METRIC, G, CH, false>;
    const size_t smem = (size_t)BUILD_WARPS * (BUILD_TMAX * 16 + (p.M + 32) * 4);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<num_sms * 4, BUILD_WARPS * 32, smem, s>>>(p);
    return cudaGetLastError();
  }
}
template <int DT, int METRIC>
static cudaError_t build_gc(const fnb_index* ix, const BuildParams& p, int num_sms, cudaStream_t s) {
  const int ch = fnb_chunks_per_lane(ix->nchunks, ix->G);
  if (ix->G == 4) {
    if (ch <= 1) return launch_build<DT, METRIC, 4, 1>(p, num_sms, s);
    return launch_build<DT, METRIC, 4, 2>(p, num_sms, s);
  }
  if (ix->G == 8) {
    switch (ch) {
      case 1: return launch_build<DT, METRIC, 8, 1>(p, num_sms, s);
      case 2: return launch_build<DT, METRIC, 8, 2>(p, num_sms, s);
      case 3: return launch_build<DT, METRIC, 8, 3>(p, num_sms, s);
      default: return launch_build<DT, METRIC, 8, 4>(p, num_sms, s);
    }
  }
  if (ch <= 2) return launch_build<DT, METRIC, 32, 2>(p, num_sms, s);
  if (ch <= 4) return launch_build<DT, METRIC, 32, 4>(p, num_sms, s);
  if (ch <= 8) return launch_build<DT, METRIC, 32, 8>(p, num_sms, s);
  return launch_build<DT, METRIC, 32, 16>(p, num_sms, s);
}
static cudaError_t launch_build_any(const fnb_index* ix, const BuildParams& p, int num_sms, cudaStream_t s) {
  const bool ip = ix->h.metric == FNB_METRIC_IP;
  switch (ix->h.data_type) {
    case FNB_DTYPE_FLOAT32: return ip ? build_gc<DT_F32, M_IP>(ix, p, num_sms, s) : build_gc<DT_F32, M_L2>(ix, p, num_sms, s);
    case FNB_DTYPE_UINT8: return ip ? build_gc<DT_U8, M_IP>(ix, p, num_sms, s) : build_gc<DT_U8, M_L2>(ix, p, num_sms, s);
    default: return ip ? build_gc<DT_I8, M_IP>(ix, p, num_sms, s) : build_gc<DT_I8, M_L2>(ix, p, num_sms, s);
  }
}

}  // namespace fnb

using namespace fnb;

#define B_CU(call)                                                                                         \
  do {                                                                                                     \
    cudaError_t e__ = (call);                                                                              \
    if (e__ != cudaSuccess) {                                                                              \
      for (void* ptr__ : tmp) cudaFree(ptr__);                                                             \
      cudaSetDevice(prev);                                                                                 \
      return fail(FNB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    }                                                                                                      \
  } while (0)

// Vectors (host, dense [n][dim]) -> padded rows [cur_nodes, cur_nodes + n), labels, and optionally all-self-loop link
// rows (Index::allocateNode, Index.h:262-272).  Does not change cur_nodes.  The caller holds ix->mu.
int upload_new_rows_locked(fnb_index* ix, const void* vectors, const int32_t* labels, int32_t label_base, int64_t n,
                           bool init_links) {
  Header& h = ix->h;
  Replica& r = ix->replicas[0];
  int prev = 0;
  cudaGetDevice(&prev);
  std::vector<void*> tmp;
  B_CU(cudaSetDevice(r.device));
  cudaStream_t s = r.stream;
  const uint32_t first = (uint32_t)h.cur_nodes;
  const size_t rowb = (size_t)ix->stride * FNB_CHUNK_BYTES;
  unsigned char* d_src = nullptr;
  const size_t chunk_rows = std::max<size_t>(1, (size_t)(256u << 20) / h.data_size);  // 256 MB staging
  B_CU(cudaMalloc(&d_src, std::min<size_t>(chunk_rows, (size_t)n) * h.data_size));
  tmp.push_back(d_src);
  for (size_t lo = 0; lo < (size_t)n; lo += chunk_rows) {
    const size_t cnt = std::min(chunk_rows, (size_t)n - lo);
    B_CU(cudaMemcpyAsync(d_src, (const unsigned char*)vectors + lo * h.data_size, cnt * h.data_size, cudaMemcpyHostToDevice, s));
    build_pad_rows_kernel<<<r.num_sms * 8, 256, 0, s>>>(d_src, (uint32_t)h.data_size,
                                                       reinterpret_cast<unsigned char*>(r.vec) + (first + lo) * rowb,
                                                       (uint32_t)rowb, cnt);
    B_CU(cudaGetLastError());
    B_CU(cudaStreamSynchronize(s));  // the pageable source buffer is reused by the caller's next chunk
  }
  std::vector<int32_t> lab;
  if (!labels) {
    lab.resize(n);
    for (int64_t i = 0; i < n; i++) lab[i] = label_base + (int32_t)i;
    labels = lab.data();
  }
  B_CU(cudaMemcpy(r.labels + first, labels, (size_t)n * 4, cudaMemcpyHostToDevice));
  if (init_links) {
    build_init_rows_kernel<<<(unsigned)(((size_t)n * h.M + 255) / 256), 256, 0, s>>>(r.adj, nullptr, first, (uint32_t)n, (uint32_t)h.M);
    B_CU(cudaGetLastError());
    B_CU(cudaStreamSynchronize(s));
  }
  for (void* ptr : tmp) cudaFree(ptr);
  cudaSetDevice(prev);
  return FNB_OK;
}

extern "C" {

int fnb_index_create(int metric, int data_type, uint64_t dim, uint64_t max_node_count, uint64_t max_edges_per_node,
                     int device, fnb_index** out) {
  if (!out) return fail(FNB_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (metric != FNB_METRIC_L2 && metric != FNB_METRIC_IP) return fail(FNB_ERR_INVALID_ARG, "unknown metric %d", metric);
  const uint64_t es = data_type == FNB_DTYPE_FLOAT32 ? 4 : (data_type == FNB_DTYPE_UINT8 || data_type == FNB_DTYPE_INT8) ? 1 : 0;
  if (!es) return fail(FNB_ERR_INVALID_ARG, "unsupported data_type %d", data_type);
  if (dim == 0 || max_edges_per_node == 0 || max_node_count == 0)
    return fail(FNB_ERR_INVALID_ARG, "dim, dataset_size and max_edges_per_node must be positive");
  if (max_node_count >= (1ull << 31)) return fail(FNB_ERR_UNSUPPORTED, "more than 2^31 nodes");
  if (max_edges_per_node > BUILD_TMAX / 2) return fail(FNB_ERR_UNSUPPORTED, "construction supports max_edges_per_node <= %d", BUILD_TMAX / 2);
  if (fnb_nchunks(dim * es) > FNB_MAX_CHUNKS) return fail(FNB_ERR_UNSUPPORTED, "vector too long for the kernels");
  // An empty index is a header-only image of the reference's file: build it through the loader so that every
  // invariant is checked in one place, then grow the device arrays to max_node_count.
  std::vector<unsigned char> img(FNB_HEADER_BYTES, 0);
  const int32_t dt = data_type;
  const uint64_t ds = dim * es, ns = ds + 4 * max_edges_per_node + 4;
  const uint64_t v[7] = {max_edges_per_node, ds, ns, 0 /* max nodes: patched below */, 0, dim, ds};
  memcpy(img.data(), &dt, 4);
  memcpy(img.data() + 4, v, 56);
  const int devs[1] = {device};
  int rc = fnb_index_from_memory(img.data(), img.size(), metric, data_type, device >= 0 ? devs : nullptr, device >= 0 ? 1 : 0, out);
  if (rc != FNB_OK) return rc;
  fnb_index* ix = *out;
  rc = fnb_index_reserve(ix, max_node_count);
  if (rc != FNB_OK) {
    fnb_index_free(ix);
    *out = nullptr;
  }
  return rc;
}

// caller holds the exclusive lock
static int reserve_locked(fnb_index* ix, uint64_t max_node_count) {
  if (ix->replicas.size() != 1) return fail(FNB_ERR_UNSUPPORTED, "construction works on a single-device index");
  if (max_node_count < ix->h.cur_nodes) return fail(FNB_ERR_INVALID_ARG, "cannot shrink below the current node count");
  if (max_node_count >= (1ull << 31)) return fail(FNB_ERR_UNSUPPORTED, "more than 2^31 nodes");
  Replica& r = ix->replicas[0];
  Header& h = ix->h;
  if (max_node_count <= r.capacity) {
    h.max_nodes = std::max<uint64_t>(h.max_nodes, max_node_count);
    return FNB_OK;
  }
  int prev = 0;
  cudaGetDevice(&prev);
  std::vector<void*> tmp;
  B_CU(cudaSetDevice(r.device));
  const size_t rowb = (size_t)ix->stride * FNB_CHUNK_BYTES;
  uint4* vec = nullptr;
  uint32_t* adj = nullptr;
  int32_t* labels = nullptr;
  B_CU(cudaMalloc(&vec, max_node_count * rowb));
  tmp.push_back(vec);
  B_CU(cudaMalloc(&adj, max_node_count * h.M * 4));
  tmp.push_back(adj);
  B_CU(cudaMalloc(&labels, max_node_count * 4));
  tmp.push_back(labels);
  if (h.cur_nodes) {
    B_CU(cudaMemcpy(vec, r.vec, h.cur_nodes * rowb, cudaMemcpyDeviceToDevice));
    B_CU(cudaMemcpy(adj, r.adj, h.cur_nodes * h.M * 4, cudaMemcpyDeviceToDevice));
    B_CU(cudaMemcpy(labels, r.labels, h.cur_nodes * 4, cudaMemcpyDeviceToDevice));
  }
  cudaFree(r.vec);
  cudaFree(r.adj);
  cudaFree(r.labels);
  r.vec = vec;
  r.adj = adj;
  r.labels = labels;
  r.capacity = max_node_count;
  r.device_bytes = max_node_count * (rowb + h.M * 4 + 4);
  h.max_nodes = max_node_count;
  cudaSetDevice(prev);
  return FNB_OK;
}

int fnb_index_reserve(fnb_index* ix, uint64_t max_node_count) {
  if (!ix) return fail(FNB_ERR_INVALID_ARG, "index is NULL");
  fnb::ExclusiveLock lock(ix->mu);
  fnb::quiesce(ix);  // an asynchronous search may still be reading the arrays that are about to be replaced
  return reserve_locked(ix, max_node_count);
}

}  // extern "C"

// Construction scratch kept on the index between fnb_index_add calls (reference-style loops that add one vector or a
// small batch at a time must not pay a dozen cudaMalloc / cudaFree and a recount of every node's degree per call).
struct fnb_build_scratch {
  uint32_t *deg = nullptr, *ovf_head = nullptr;  // [node_cap]
  uint64_t node_cap = 0;
  uint64_t deg_nodes = 0;  // deg[] is current for nodes [0, deg_nodes) — reset by anything else that edits link rows
  uint32_t *ovf_next = nullptr, *ovf_src = nullptr, *dirty = nullptr;  // [ovf_cap]
  uint64_t ovf_cap = 0;
  int32_t* cand_id = nullptr;  // [cand_cap]
  float* cand_dist = nullptr;
  uint64_t cand_cap = 0;
  unsigned int* counters = nullptr;  // 64 B
  cudaEvent_t e0 = nullptr, e1 = nullptr;
};

void fnb_build_scratch_free(fnb_build_scratch* b) {
  if (!b) return;
  cudaFree(b->deg);
  cudaFree(b->ovf_head);
  cudaFree(b->ovf_next);
  cudaFree(b->ovf_src);
  cudaFree(b->dirty);
  cudaFree(b->cand_id);
  cudaFree(b->cand_dist);
  cudaFree(b->counters);
  if (b->e0) cudaEventDestroy(b->e0);
  if (b->e1) cudaEventDestroy(b->e1);
  delete b;
}

void fnb_index_mutated(fnb_index* ix, bool links, bool labels) {
  if (links && ix->build) ix->build->deg_nodes = 0;
  if (labels && ix->label_map) {
    fnb_label_map_free(ix->label_map);
    ix->label_map = nullptr;
  }
}

#define S_CU(call)                                                                                            \
  do {                                                                                                        \
    cudaError_t e__ = (call);                                                                                 \
    if (e__ != cudaSuccess)                                                                                   \
      return fail(FNB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

template <typename T>
static int grow(T** ptr, uint64_t* cap, uint64_t want, size_t elem, T** twin = nullptr, T** twin2 = nullptr) {
  if (want <= *cap) return FNB_OK;
  const uint64_t n = std::max<uint64_t>(want, *cap + *cap / 2);
  for (T** q : {ptr, twin, twin2}) {
    if (!q) continue;
    cudaFree(*q);
    *q = nullptr;
  }
  *cap = 0;
  for (T** q : {ptr, twin, twin2}) {
    if (!q) continue;
    S_CU(cudaMalloc((void**)q, n * elem));
  }
  *cap = n;
  return FNB_OK;
}

// The body of fnb_index_add; the caller holds the exclusive lock, has validated the arguments and restores the
// current device.  On a CUDA failure inside the batch loop the index is left with the nodes of the batches that
// completed (cur_num_nodes says how many): every link still points at a live node.
static int add_locked(fnb_index* ix, const void* vectors, const int32_t* labels, int64_t n, int ef_construction,
                      int num_initializations, fnb_build_stats* stats) {
  Header& h = ix->h;
  Replica& r = ix->replicas[0];
  const uint32_t M = (uint32_t)h.M, Msel = std::max(M / 2u, 1u);
  const uint32_t Kc = (uint32_t)ef_construction;
  S_CU(cudaSetDevice(r.device));
  cudaStream_t s = r.stream;
  if (!ix->build) ix->build = new fnb_build_scratch();
  fnb_build_scratch& B = *ix->build;
  if (!B.e0) S_CU(cudaEventCreate(&B.e0));
  if (!B.e1) S_CU(cudaEventCreate(&B.e1));
  if (!B.counters) S_CU(cudaMalloc((void**)&B.counters, 64));
  S_CU(cudaEventRecord(B.e0, s));
  const uint32_t first = (uint32_t)h.cur_nodes, total = first + (uint32_t)n;
  const uint32_t max_b = (uint32_t)std::max(1, getenv("FNB_BUILD_BATCH") ? atoi(getenv("FNB_BUILD_BATCH")) : 16384);
  const uint32_t bmax = (uint32_t)std::min<uint64_t>(max_b, (uint64_t)n);  // no batch of this call is larger
  const uint32_t ovf_cap = bmax * Msel;

  // ---- upload: vectors (padded rows), labels ----
  int rc = upload_new_rows_locked(ix, vectors, labels, 0, n, /*init_links=*/false);  // labels NULL: 0 .. n-1 (bindings.cpp:84-86)
  if (rc != FNB_OK) return rc;
  // ---- scratch: grown, never shrunk ----
  {
    const uint64_t had = B.node_cap;
    rc = grow(&B.deg, &B.node_cap, r.capacity, 4, &B.ovf_head);
    if (rc != FNB_OK) return rc;
    if (B.node_cap != had) B.deg_nodes = 0;
    rc = grow(&B.ovf_next, &B.ovf_cap, ovf_cap, 4, &B.ovf_src, &B.dirty);
    if (rc != FNB_OK) return rc;
    uint64_t cc = B.cand_cap;
    rc = grow(&B.cand_id, &cc, (uint64_t)bmax * Kc, 4);
    if (rc != FNB_OK) return rc;
    rc = grow(&B.cand_dist, &B.cand_cap, (uint64_t)bmax * Kc, 4);
    if (rc != FNB_OK) return rc;
  }
  S_CU(cudaMemsetAsync(B.ovf_head, 0, (size_t)total * 4, s));
  S_CU(cudaMemsetAsync(B.counters, 0, 64, s));
  if (B.deg_nodes != first && first) {  // degrees of the existing rows are not known (loaded / re-ordered / imported)
    build_count_degree_kernel<<<(first + 255) / 256, 256, 0, s>>>(r.adj, B.deg, first, M, B.counters + 8);
    S_CU(cudaGetLastError());
    unsigned int unpacked = 0;
    S_CU(cudaMemcpyAsync(&unpacked, B.counters + 8, 4, cudaMemcpyDeviceToHost, s));
    S_CU(cudaStreamSynchronize(s));
    if (unpacked)
      return fail(FNB_ERR_UNSUPPORTED, "%u nodes have link rows with self-loops before real links; cannot append", unpacked);
  }
  B.deg_nodes = 0;  // until this call has finished
  build_init_rows_kernel<<<(unsigned)(((size_t)n * M + 255) / 256), 256, 0, s>>>(r.adj, B.deg, first, (uint32_t)n, M);
  S_CU(cudaGetLastError());

  // ---- batched insertion ----
  SearchParams sp;
  int64_t n_batches = 0;
  uint32_t done = first;
  if (done == 0) done = 1;  // the first node has nothing to link to (Index.h:366-368)
  auto enqueue_batch = [&](uint32_t b) -> int {
    // the graph the batch searches is the first `done` nodes; cur_num_nodes itself moves only when a batch is complete
    int prc = plan_search(ix, b, (int)Kc, (int)Kc, num_initializations, &sp, 0, done, /*allow_latency_variant=*/false);
    if (prc != FNB_OK) return prc;
    sp.vec = r.vec;
    sp.adj = r.adj;
    sp.labels = nullptr;  // node ids, not labels
    sp.queries = r.vec + (size_t)done * ix->stride;
    sp.query_pitch_chunks = ix->stride;
    sp.out_dist = B.cand_dist;
    sp.out_label = B.cand_id;
    sp.counter = r.counter;
    sp.totals = r.totals;
    S_CU(cudaMemsetAsync(r.counter, 0, 128, s));
    S_CU(dispatch_search(ix, sp, r.num_sms, s));
    BuildParams bp;
    bp.sp = sp;
    bp.adj = r.adj;
    bp.deg = B.deg;
    bp.cand_id = B.cand_id;
    bp.cand_dist = B.cand_dist;
    bp.ovf_head = B.ovf_head;
    bp.ovf_next = B.ovf_next;
    bp.ovf_src = B.ovf_src;
    bp.dirty = B.dirty;
    bp.counters = B.counters;
    bp.M = M;
    bp.Msel = Msel;
    bp.Kc = Kc;
    bp.first = done;
    bp.b = b;
    bp.ovf_cap = ovf_cap;
    S_CU(cudaMemsetAsync(B.counters, 0, 8, s));      // overflow entries, dirty nodes
    S_CU(cudaMemsetAsync(B.counters + 3, 0, 4, s));  // prune work counter
    S_CU(launch_build_any(ix, bp, r.num_sms, s));
    return FNB_OK;
  };
  while (done < total) {
    // nodes of one batch cannot link to each other: keep a batch small against the graph it is inserted into
    // (1/16 while the graph is small and every link matters, 1/8 afterwards)
    const uint32_t b = std::min({max_b, total - done, std::max(1u, done < 65536u ? done / 16u : done / 8u)});
    rc = enqueue_batch(b);
    if (rc != FNB_OK) {
      // keep what the completed batches built; the rows beyond are uploaded but not part of the graph
      const std::string keep = g_last_error;
      if (cudaStreamSynchronize(s) == cudaSuccess) h.cur_nodes = done;
      g_last_error = keep;
      return rc;
    }
    done += b;
    n_batches++;
  }
  unsigned int hc[4] = {0, 0, 0, 0};
  S_CU(cudaMemcpyAsync(hc, B.counters, 16, cudaMemcpyDeviceToHost, s));
  S_CU(cudaEventRecord(B.e1, s));
  S_CU(cudaStreamSynchronize(s));
  h.cur_nodes = total;
  B.deg_nodes = total;
  fnb_index_mutated(ix, false, true);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, B.e0, B.e1);
  if (stats) {
    stats->n_added = n;
    stats->n_batches = n_batches;
    stats->n_dropped_backlinks = hc[2];
    stats->device_ms = ms;
  }
  return FNB_OK;
}

extern "C" {

int fnb_index_add(fnb_index* ix, const void* vectors, const int32_t* labels, int64_t n, int ef_construction,
                  int num_initializations, fnb_build_stats* stats) {
  if (!ix) return fail(FNB_ERR_INVALID_ARG, "index is NULL");
  if (stats) memset(stats, 0, sizeof(*stats));
  if (n < 0) return fail(FNB_ERR_INVALID_ARG, "negative vector count");
  if (num_initializations <= 0) return fail(FNB_ERR_INVALID_ARG, "num_initializations must be greater than 0.");  // Index.h:304
  if (ef_construction <= 0) return fail(FNB_ERR_INVALID_ARG, "ef_construction must be positive");
  if (n == 0) return FNB_OK;
  if (!vectors) return fail(FNB_ERR_INVALID_ARG, "vectors is NULL");
  if (ix->replicas.size() != 1) return fail(FNB_ERR_UNSUPPORTED, "construction works on a single-device index");
  fnb::ExclusiveLock lock(ix->mu);
  fnb::quiesce(ix);
  Header& h = ix->h;
  Replica& r = ix->replicas[0];
  if (h.cur_nodes + (uint64_t)n > h.max_nodes)
    return fail(FNB_ERR_INVALID_ARG, "Maximum number of nodes reached. Consider increasing the `max_node_count` parameter to "
                                     "create a larger index.");  // Index.h:356-361
  if ((uint32_t)h.M > BUILD_TMAX / 2) return fail(FNB_ERR_UNSUPPORTED, "construction supports max_edges_per_node <= %d", BUILD_TMAX / 2);
  int prev = 0;
  cudaGetDevice(&prev);
  int rc = FNB_OK;
  // a loaded index holds exactly cur_num_nodes rows although its header allows max_node_count: make room
  if (h.cur_nodes + (uint64_t)n > r.capacity) rc = reserve_locked(ix, h.max_nodes);
  if (rc == FNB_OK) rc = add_locked(ix, vectors, labels, n, ef_construction, num_initializations, stats);
  cudaSetDevice(prev);
  return rc;
}

}  // extern "C"
