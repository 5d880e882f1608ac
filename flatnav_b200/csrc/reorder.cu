// Graph re-ordering and link import for a loaded index.
//
// Replaces, of the reference,
//   Index::doGraphReordering / reorderGOrder / reorderRCM   include/flatnav/index/Index.h:412-440
//   util::gOrder                                            include/flatnav/util/Reordering.h:26-117
//   util::GorderPriorityQueue                               include/flatnav/util/GorderPriorityQueue.h:13-112
//   util::rcmOrder                                          include/flatnav/util/Reordering.h:119-199
//   Index::relabel / swapNodes                              Index.h:872-926, 575-592
//   Index::getGraphOutdegreeTable                           Index.h:240-251
//   Index::allocateNode / buildGraphLinks                   Index.h:262-272, 187-238
//
// Split of the work.  Both orderings are *sequential greedy* algorithms whose output (and therefore the file a
// re-ordered index is saved to) depends on the exact order of their queue operations; they run on the host, as in
// the reference, but with different data structures (below).  Everything that touches the index data — the
// extraction of the link table, the relabelling of all N*M links and the physical re-layout of N vector rows,
// link rows and labels — runs on the GPU, out of place, one warp per node, on every replica.
//
//   gorder  The reference keeps (key, priority) pairs in a vector sorted by priority, finds the block boundary of a
//           priority with std::upper_bound / std::lower_bound and looks keys up in an unordered_map: O(log N) plus
//           a hash probe per increment/decrement, ~2200 of them per node at M=32.  Here: flat arrays pos[key],
//           prio[key], key_at[slot] and a lazily clamped boundary table lo[p] = number of entries with priority
//           < p.  "rightmost entry of my priority" is min(lo[p+1], size) - 1 and "leftmost" is min(lo[p], size),
//           both O(1); the clamp makes pop() O(1) too (proof in the comment of GorderQueue).  The swaps performed
//           are the reference's, so the permutation is identical.
//   rcm     as the reference (degree-sorted BFS, std::sort with the same comparator on the same sequences — the
//           order among equal degrees is whatever std::sort gives, in the reference and here), over CSR arrays
//           instead of vector<vector>.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <queue>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "../../include/flatnav_b200.h"
#include "fnb_internal.h"

namespace fnb {

#define R_CU(call)                                                                                         \
  do {                                                                                                     \
    cudaError_t e__ = (call);                                                                              \
    if (e__ != cudaSuccess) {                                                                              \
      for (void* ptr__ : tmp) cudaFree(ptr__);                                                             \
      cudaSetDevice(prev);                                                                                 \
      return fail(FNB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    }                                                                                                      \
  } while (0)

// ---- device side ---------------------------------------------------------------------------------------
// new[P[n]] = old[n] for the vector row, the label and the link row, whose entries are mapped through P as well
// (Index.h:874-880 rewires, :893-918 moves the nodes along the cycles of P with swapNodes; the net effect is this
// gather).  A self-loop n -> n becomes P[n] -> P[n], so unused slots stay unused.
__global__ void relabel_kernel(const uint4* __restrict__ vec, const uint32_t* __restrict__ adj,
                               const int32_t* __restrict__ labels, const uint32_t* __restrict__ perm, uint32_t n_nodes,
                               uint32_t M, uint32_t stride, uint4* __restrict__ vec_out, uint32_t* __restrict__ adj_out,
                               int32_t* __restrict__ labels_out) {
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (uint64_t n = (((uint64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; n < n_nodes; n += warps) {
    const uint64_t d = perm[n];
    for (uint32_t c = lane; c < stride; c += 32) vec_out[d * stride + c] = vec[n * stride + c];
    for (uint32_t j = lane; j < M; j += 32) adj_out[d * M + j] = perm[adj[n * M + j]];
    if (lane == 0) labels_out[d] = labels[n];
  }
}

// P must be a permutation of [0, n): every value in range, no value twice.
__global__ void check_perm_kernel(const uint32_t* __restrict__ perm, uint32_t n, uint32_t* __restrict__ seen,
                                  unsigned int* __restrict__ bad) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = perm[i];
  if (v >= n || (atomicOr(seen + (v >> 5), 1u << (v & 31)) >> (v & 31)) & 1u) atomicAdd(bad, 1u);
}

// buildGraphLinks (Index.h:219-234): edges (u, v) in file order; v takes the first slot of u's row that still
// points at u.  Edges of one source node must keep their order; different source nodes are independent, so the
// host groups the edges by source (stable) and one thread fills one row.
__global__ void fill_links_kernel(uint32_t* __restrict__ adj, uint32_t M, const uint32_t* __restrict__ src_start,
                                  const uint32_t* __restrict__ dst, uint32_t n_nodes) {
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_nodes) return;
  uint32_t* row = adj + (size_t)u * M;
  uint32_t slot = 0;
  for (uint32_t e = src_start[u]; e < src_start[u + 1]; e++) {
    while (slot < M && row[slot] != u) slot++;
    if (slot >= M) break;  // row full: the reference silently drops the edge
    row[slot] = dst[e];
    // a self-edge (v == u) leaves the slot "available" (it still points at u): the reference re-uses it for the
    // next edge, and so does this scan because `slot` is not advanced past a value equal to u
  }
}

// ---- host side: the two orderings -------------------------------------------------------------------------
struct Csr {
  std::vector<uint64_t> start;  // [n + 1]
  std::vector<uint32_t> item;
};

// getGraphOutdegreeTable (Index.h:240-251): the links of every node that are not self-loops, in slot order.
static Csr out_table(const uint32_t* links, uint32_t n, uint32_t M) {
  Csr t;
  t.start.assign((size_t)n + 1, 0);
  for (uint32_t v = 0; v < n; v++) {
    uint32_t d = 0;
    for (uint32_t j = 0; j < M; j++) d += links[(size_t)v * M + j] != v;
    t.start[v + 1] = t.start[v] + d;
  }
  t.item.resize(t.start[n]);
  for (uint32_t v = 0; v < n; v++) {
    uint64_t w = t.start[v];
    for (uint32_t j = 0; j < M; j++) {
      const uint32_t x = links[(size_t)v * M + j];
      if (x != v) t.item[w++] = x;
    }
  }
  return t;
}

// Reordering.h:56-62: in-edges listed by ascending source node, one entry per occurrence.
static Csr in_table(const Csr& out, uint32_t n) {
  Csr t;
  t.start.assign((size_t)n + 1, 0);
  for (uint32_t x : out.item) t.start[(size_t)x + 1]++;
  for (uint32_t v = 0; v < n; v++) t.start[v + 1] += t.start[v];
  t.item.resize(out.item.size());
  std::vector<uint64_t> w(t.start.begin(), t.start.end() - 1);
  for (uint32_t v = 0; v < n; v++)
    for (uint64_t e = out.start[v]; e < out.start[v + 1]; e++) t.item[w[out.item[e]]++] = v;
  return t;
}

// GorderPriorityQueue with O(1) operations.  Invariants (the reference's, GorderPriorityQueue.h:59-103): slots
// [0, size) hold the live keys in ascending priority; increment swaps the key with the RIGHTMOST entry of its
// priority, decrement with the LEFTMOST, pop removes the last slot.
// lo[p] is meant to be the number of live entries with priority < p.  It is kept lazily: the true value is always
// min(lo[p], size).  increment / decrement touch exactly one boundary and write the clamped value back; pop of an
// entry of priority pm lowers the true value of every p > pm from size to size - 1, which the clamp delivers
// without touching the table (for p <= pm the popped entry was not counted, and lo[p] <= size - 1 already).
struct GorderQueue {
  struct State {
    int32_t pos;   // key -> slot, -1 once popped
    int32_t prio;  // key -> priority
  };
  std::vector<State> st;         // one cache line access per key for both fields
  std::vector<uint32_t> key_at;  // slot -> key
  std::vector<uint32_t> lo;      // boundary table, indexed by priority + kOff
  uint32_t size, n;
  static constexpr int32_t kOff = 4;  // priorities never go negative (every decrement undoes an increment of a
                                      // key that was already live then); the offset is a safety margin only
  explicit GorderQueue(uint32_t n_) : st(n_), key_at(n_), lo(64, n_), size(n_), n(n_) {
    for (uint32_t i = 0; i < n_; i++) {
      st[i].pos = (int32_t)i;
      st[i].prio = 0;
      key_at[i] = i;
    }
    for (int32_t p = 0; p <= kOff; p++) lo[p] = 0;  // nothing has a priority below 0
  }
  inline void prefetch(uint32_t key) const { __builtin_prefetch(&st[key], 1, 1); }
  inline uint32_t bound(int32_t p) {  // clamped lo[p]
    const size_t i = (size_t)(p + kOff);
    if (i >= lo.size()) lo.resize(std::max(lo.size() * 2, i + 1), n);
    return std::min(lo[i], size);
  }
  inline void swap_slots(uint32_t a, uint32_t b) {
    const uint32_t ka = key_at[a], kb = key_at[b];
    key_at[a] = kb;
    key_at[b] = ka;
    st[kb].pos = (int32_t)a;
    st[ka].pos = (int32_t)b;
  }
  inline bool increment(uint32_t key) {
    State& k = st[key];
    const int32_t s = k.pos;
    if (s < 0) return true;
    const int32_t p = k.prio;
    const uint32_t idx = bound(p + 1) - 1;  // upper_bound - 1: rightmost entry of priority p
    k.prio = p + 1;
    if ((uint32_t)s != idx) swap_slots((uint32_t)s, idx);
    lo[(size_t)(p + 1 + kOff)] = idx;
    return true;
  }
  inline bool decrement(uint32_t key) {
    State& k = st[key];
    const int32_t s = k.pos;
    if (s < 0) return true;
    const int32_t p = k.prio;
    if (p + kOff <= 0) return false;
    const uint32_t idx = bound(p);  // lower_bound: leftmost entry of priority p
    k.prio = p - 1;
    if ((uint32_t)s != idx) swap_slots((uint32_t)s, idx);
    lo[(size_t)(p + kOff)] = idx + 1;
    return true;
  }
  inline uint32_t pop() {
    const uint32_t key = key_at[--size];
    st[key].pos = -1;
    return key;
  }
};

// util::gOrder (Reordering.h:26-117).  Returns P with P[old id] = new id.
static int gorder_perm(const uint32_t* links, uint32_t n, uint32_t M, int w, std::vector<uint32_t>* perm) {
  perm->assign(n, 0);
  if (n == 0) return FNB_OK;
  const Csr out = out_table(links, n, M);
  const Csr in = in_table(out, n);
  GorderQueue q(n);
  std::vector<uint32_t> order(n, 0);  // the reference's P: order[i] = node placed i-th
  bool ok = true;
  q.increment(0);  // seed node 0 (Reordering.h:67-69)
  order[0] = q.pop();
  for (int64_t i = 1; i < (int64_t)n; i++) {
    const uint32_t ve = order[i - 1];
    // The order of the queue operations is the reference's; prefetches only warm the per-key state of the keys a few
    // operations ahead (the sequence is a random walk over ~2200 keys per placed node).
    constexpr uint64_t kAhead = 12;
    const uint64_t n_items = out.item.size();
    for (uint64_t e = out.start[ve]; e < out.start[ve + 1]; e++) q.increment(out.item[e]);
    for (uint64_t e = in.start[ve]; e < in.start[ve + 1]; e++) {
      const uint32_t u = in.item[e];
      if (e + 1 < in.start[ve + 1]) __builtin_prefetch(&out.item[out.start[in.item[e + 1]]], 0, 1);
      q.increment(u);
      const uint64_t f1 = out.start[u + 1];
      for (uint64_t f = out.start[u]; f < f1; f++) {
        if (f + kAhead < n_items) q.prefetch(out.item[f + kAhead]);
        q.increment(out.item[f]);
      }
    }
    if (i > (int64_t)w + 1) {
      const uint32_t vb = order[i - w - 1];
      for (uint64_t e = out.start[vb]; e < out.start[vb + 1]; e++) ok &= q.decrement(out.item[e]);
      for (uint64_t e = in.start[vb]; e < in.start[vb + 1]; e++) {
        const uint32_t u = in.item[e];
        if (e + 1 < in.start[vb + 1]) __builtin_prefetch(&out.item[out.start[in.item[e + 1]]], 0, 1);
        ok &= q.decrement(u);
        const uint64_t f1 = out.start[u + 1];
        for (uint64_t f = out.start[u]; f < f1; f++) {
          if (f + kAhead < n_items) q.prefetch(out.item[f + kAhead]);
          ok &= q.decrement(out.item[f]);
        }
      }
    }
    order[i] = q.pop();
  }
  if (!ok) return fail(FNB_ERR_UNSUPPORTED, "gorder: a priority fell below the supported range");
  for (uint32_t i = 0; i < n; i++) (*perm)[order[i]] = i;
  return FNB_OK;
}

// util::rcmOrder (Reordering.h:119-199).  Returns P with P[old id] = new id.
static int rcm_perm(const uint32_t* links, uint32_t n, uint32_t M, std::vector<uint32_t>* perm) {
  perm->assign(n, 0);
  if (n == 0) return FNB_OK;
  const Csr out = out_table(links, n, M);
  typedef std::pair<uint32_t, int> node_deg;
  auto by_degree = [](const node_deg& a, const node_deg& b) { return a.second < b.second; };
  std::vector<int> degrees(n);
  std::vector<node_deg> sorted_nodes(n);
  for (uint32_t v = 0; v < n; v++) {
    degrees[v] = (int)(out.start[v + 1] - out.start[v]);
    sorted_nodes[v] = node_deg(v, degrees[v]);
  }
  std::sort(sorted_nodes.begin(), sorted_nodes.end(), by_degree);
  std::vector<uint32_t> order;
  order.reserve(n);
  std::vector<unsigned char> visited(n, 0);
  std::vector<node_deg> nb;
  std::queue<uint32_t> fifo;
  auto push_neighbours = [&](uint32_t v) {  // neighbours by ascending degree (Reordering.h:146-160, 172-186)
    nb.clear();
    for (uint64_t e = out.start[v]; e < out.start[v + 1]; e++) nb.push_back(node_deg(out.item[e], degrees[out.item[e]]));
    std::sort(nb.begin(), nb.end(), by_degree);
    for (const node_deg& x : nb) fifo.push(x.first);
  };
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t root = sorted_nodes[i].first;
    if (visited[root]) continue;
    order.push_back(root);
    visited[root] = 1;
    push_neighbours(root);
    while (!fifo.empty()) {
      const uint32_t c = fifo.front();
      fifo.pop();
      if (visited[c]) continue;
      order.push_back(c);
      visited[c] = 1;
      push_neighbours(c);
    }
  }
  std::reverse(order.begin(), order.end());
  for (uint32_t i = 0; i < n; i++) (*perm)[order[i]] = i;
  return FNB_OK;
}

// ---- host side: device plumbing -----------------------------------------------------------------------------
static int download_links(const fnb_index* ix, std::vector<uint32_t>* links) {
  const Header& h = ix->h;
  const Replica& r = ix->replicas[0];
  links->resize((size_t)h.cur_nodes * h.M);
  if (!h.cur_nodes) return FNB_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  std::vector<void*> tmp;
  R_CU(cudaSetDevice(r.device));
  R_CU(cudaStreamSynchronize(r.stream));
  R_CU(cudaMemcpy(links->data(), r.adj, links->size() * 4, cudaMemcpyDeviceToHost));
  cudaSetDevice(prev);
  return FNB_OK;
}

// Two phases, so that replicas can never end up permuted differently: every replica first builds its re-laid-out
// arrays out of place; only when all of them succeeded are the arrays swapped in.  A failure on any replica frees what
// was built and leaves the whole index as it was.
static int apply_perm(fnb_index* ix, const uint32_t* perm) {
  const Header& h = ix->h;
  const uint32_t n = (uint32_t)h.cur_nodes;
  if (!n) return FNB_OK;
  struct NewArrays {
    int device = -1;
    uint4* vec = nullptr;
    uint32_t* adj = nullptr;
    int32_t* labels = nullptr;
  };
  std::vector<NewArrays> built(ix->replicas.size());
  int prev = 0;
  cudaGetDevice(&prev);
  auto build_one = [&](Replica& r, NewArrays& out) -> int {
    std::vector<void*> tmp;  // R_CU frees these on failure
    R_CU(cudaSetDevice(r.device));
    out.device = r.device;
    cudaStream_t s = r.stream;
    const size_t rowb = (size_t)ix->stride * FNB_CHUNK_BYTES;
    uint32_t *d_perm = nullptr, *d_seen = nullptr;
    unsigned int* d_bad = nullptr;
    R_CU(cudaMalloc(&d_perm, (size_t)n * 4));
    tmp.push_back(d_perm);
    R_CU(cudaMalloc(&d_seen, ((size_t)n / 32 + 1) * 4));
    tmp.push_back(d_seen);
    R_CU(cudaMalloc(&d_bad, 4));
    tmp.push_back(d_bad);
    R_CU(cudaMemcpyAsync(d_perm, perm, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    R_CU(cudaMemsetAsync(d_seen, 0, ((size_t)n / 32 + 1) * 4, s));
    R_CU(cudaMemsetAsync(d_bad, 0, 4, s));
    check_perm_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_perm, n, d_seen, d_bad);
    R_CU(cudaGetLastError());
    unsigned int bad = 0;
    R_CU(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, s));
    R_CU(cudaStreamSynchronize(s));
    if (bad) {
      for (void* ptr : tmp) cudaFree(ptr);
      cudaSetDevice(prev);
      return fail(FNB_ERR_INVALID_ARG, "not a permutation of [0, %u): %u entries out of range or repeated", n, bad);
    }
    // out of place into arrays of the same capacity (construction may have reserved more rows than are in use)
    R_CU(cudaMalloc(&out.vec, r.capacity * rowb));
    tmp.push_back(out.vec);
    R_CU(cudaMalloc(&out.adj, r.capacity * h.M * 4));
    tmp.push_back(out.adj);
    R_CU(cudaMalloc(&out.labels, r.capacity * 4));
    tmp.push_back(out.labels);
    relabel_kernel<<<r.num_sms * 16, 256, 0, s>>>(r.vec, r.adj, r.labels, d_perm, n, (uint32_t)h.M, ix->stride, out.vec,
                                                  out.adj, out.labels);
    R_CU(cudaGetLastError());
    R_CU(cudaStreamSynchronize(s));
    cudaFree(d_perm);
    cudaFree(d_seen);
    cudaFree(d_bad);
    return FNB_OK;
  };
  for (size_t i = 0; i < ix->replicas.size(); i++) {
    const int rc = build_one(ix->replicas[i], built[i]);
    if (rc != FNB_OK) {
      const std::string keep = g_last_error;
      built[i] = NewArrays();  // its buffers were freed by the failing step
      for (size_t k = 0; k < i; k++) {
        cudaSetDevice(built[k].device);
        cudaFree(built[k].vec);
        cudaFree(built[k].adj);
        cudaFree(built[k].labels);
      }
      cudaSetDevice(prev);
      g_last_error = keep;
      return rc;
    }
  }
  for (size_t i = 0; i < ix->replicas.size(); i++) {
    Replica& r = ix->replicas[i];
    cudaSetDevice(r.device);
    cudaFree(r.vec);
    cudaFree(r.adj);
    cudaFree(r.labels);
    r.vec = built[i].vec;
    r.adj = built[i].adj;
    r.labels = built[i].labels;
  }
  cudaSetDevice(prev);
  return FNB_OK;
}

}  // namespace fnb

using namespace fnb;

extern "C" {

int fnb_index_links(const fnb_index* ix, uint32_t* out_links) {
  if (!ix || !out_links) return fail(FNB_ERR_INVALID_ARG, "NULL argument");
  fnb::ExclusiveLock lock(ix->mu);  // download_links runs on the mutator stream
  std::vector<uint32_t> links;
  int rc = download_links(ix, &links);
  if (rc != FNB_OK) return rc;
  if (!links.empty()) memcpy(out_links, links.data(), links.size() * 4);
  return FNB_OK;
}

int fnb_index_relabel(fnb_index* ix, const uint32_t* perm) {
  if (!ix || !perm) return fail(FNB_ERR_INVALID_ARG, "NULL argument");
  fnb::ExclusiveLock lock(ix->mu);
  fnb::quiesce(ix);
  const int rc = apply_perm(ix, perm);
  fnb_index_mutated(ix, true, true);
  return rc;
}

int fnb_graph_order(const uint32_t* links, uint64_t n_nodes, uint64_t max_edges_per_node, int method, int window,
                    uint32_t* perm_out) {
  if (!perm_out || (!links && n_nodes)) return fail(FNB_ERR_INVALID_ARG, "NULL argument");
  if (method != FNB_REORDER_GORDER && method != FNB_REORDER_RCM)
    return fail(FNB_ERR_INVALID_ARG, "Invalid reordering method: %d", method);
  if (n_nodes >= (1ull << 31) || max_edges_per_node == 0 || max_edges_per_node > 65536)
    return fail(FNB_ERR_INVALID_ARG, "bad graph shape");
  if (window <= 0) window = 5;
  const uint32_t n = (uint32_t)n_nodes, M = (uint32_t)max_edges_per_node;
  for (uint64_t i = 0; i < n_nodes * max_edges_per_node; i++)
    if (links[i] >= n) return fail(FNB_ERR_INVALID_ARG, "link %u outside [0, %u)", links[i], n);
  std::vector<uint32_t> perm;
  const int rc = method == FNB_REORDER_GORDER ? gorder_perm(links, n, M, window, &perm) : rcm_perm(links, n, M, &perm);
  if (rc != FNB_OK) return rc;
  if (n) memcpy(perm_out, perm.data(), (size_t)n * 4);
  return FNB_OK;
}

int fnb_index_reorder(fnb_index* ix, int method, int window, uint32_t* perm_out) {
  if (!ix) return fail(FNB_ERR_INVALID_ARG, "index is NULL");
  if (method != FNB_REORDER_GORDER && method != FNB_REORDER_RCM)
    return fail(FNB_ERR_INVALID_ARG, "Invalid reordering method: %d", method);  // Index.h:421-423
  if (window <= 0) window = 5;                                                   // Index.h:418
  fnb::ExclusiveLock lock(ix->mu);
  fnb::quiesce(ix);
  std::vector<uint32_t> links, perm;
  int rc = download_links(ix, &links);
  if (rc != FNB_OK) return rc;
  const uint32_t n = (uint32_t)ix->h.cur_nodes, M = (uint32_t)ix->h.M;
  rc = method == FNB_REORDER_GORDER ? gorder_perm(links.data(), n, M, window, &perm) : rcm_perm(links.data(), n, M, &perm);
  if (rc != FNB_OK) return rc;
  rc = apply_perm(ix, perm.data());
  fnb_index_mutated(ix, true, true);
  if (rc != FNB_OK) return rc;
  if (perm_out && n) memcpy(perm_out, perm.data(), (size_t)n * 4);
  return FNB_OK;
}

int fnb_index_allocate_nodes(fnb_index* ix, const void* vectors, const int32_t* labels, int64_t n) {
  if (!ix) return fail(FNB_ERR_INVALID_ARG, "index is NULL");
  if (n < 0) return fail(FNB_ERR_INVALID_ARG, "negative vector count");
  if (n == 0) return FNB_OK;
  if (!vectors) return fail(FNB_ERR_INVALID_ARG, "vectors is NULL");
  if (ix->replicas.size() != 1) return fail(FNB_ERR_UNSUPPORTED, "construction works on a single-device index");
  fnb::ExclusiveLock lock(ix->mu);
  fnb::quiesce(ix);
  Header& h = ix->h;
  Replica& r = ix->replicas[0];
  if (h.cur_nodes + (uint64_t)n > h.max_nodes || h.cur_nodes + (uint64_t)n > r.capacity)
    return fail(FNB_ERR_INVALID_ARG, "Maximum number of nodes reached. Consider increasing the `max_node_count` parameter to "
                                     "create a larger index.");
  // labels NULL: cur_nodes, cur_nodes + 1, ... (PyIndex::allocateNodes numbers them with its _label_id counter)
  int rc = upload_new_rows_locked(ix, vectors, labels, (int32_t)h.cur_nodes, n, /*init_links=*/true);
  if (rc != FNB_OK) return rc;
  h.cur_nodes += (uint64_t)n;
  fnb_index_mutated(ix, true, true);
  return FNB_OK;
}

int fnb_index_build_graph_links(fnb_index* ix, const char* mtx_filename) {
  if (!ix || !mtx_filename) return fail(FNB_ERR_INVALID_ARG, "NULL argument");
  if (ix->replicas.size() != 1) return fail(FNB_ERR_UNSUPPORTED, "construction works on a single-device index");
  std::ifstream in(mtx_filename);
  if (!in.is_open()) return fail(FNB_ERR_IO, "Unable to open file for reading: %s", mtx_filename);
  fnb::ExclusiveLock lock(ix->mu);
  fnb::quiesce(ix);
  const Header& h = ix->h;
  std::string line;
  while (std::getline(in, line)) {  // skip the '%' header lines; the first other line is the size line
    if (line.empty() || line[0] != '%') break;
  }
  std::istringstream iss(line);
  long long nv = 0, ne = 0;
  iss >> nv >> nv >> ne;  // rows, columns, "edges" — which the reference requires to be M (Index.h:200-217)
  if ((uint64_t)nv != h.max_nodes)
    return fail(FNB_ERR_IO, "Number of vertices in the mtx file does not match the size allocated for the index.");
  if ((uint64_t)ne != h.M)
    return fail(FNB_ERR_IO, "Number of edges in the mtx file does not match the number of links per node.");
  const uint32_t n = (uint32_t)h.cur_nodes;
  std::vector<std::pair<uint32_t, uint32_t>> edges;
  long long u, v;
  while (in >> u >> v) {
    u--;  // Matrix Market is 1-based
    v--;
    if (u < 0 || v < 0 || u >= (long long)n || v >= (long long)n)
      return fail(FNB_ERR_FORMAT, "edge (%lld, %lld) outside the %u allocated nodes", u + 1, v + 1, n);
    edges.emplace_back((uint32_t)u, (uint32_t)v);
  }
  if (!n) return FNB_OK;
  // group by source node, keeping file order within a node
  std::vector<uint32_t> start((size_t)n + 1, 0), dst(edges.size());
  for (auto& e : edges) start[(size_t)e.first + 1]++;
  for (uint32_t i = 0; i < n; i++) start[i + 1] += start[i];
  {
    std::vector<uint32_t> w(start.begin(), start.end() - 1);
    for (auto& e : edges) dst[w[e.first]++] = e.second;
  }
  Replica& r = ix->replicas[0];
  int prev = 0;
  cudaGetDevice(&prev);
  std::vector<void*> tmp;
  R_CU(cudaSetDevice(r.device));
  uint32_t *d_start = nullptr, *d_dst = nullptr;
  R_CU(cudaMalloc(&d_start, start.size() * 4));
  tmp.push_back(d_start);
  R_CU(cudaMalloc(&d_dst, std::max<size_t>(dst.size(), 1) * 4));
  tmp.push_back(d_dst);
  R_CU(cudaMemcpyAsync(d_start, start.data(), start.size() * 4, cudaMemcpyHostToDevice, r.stream));
  if (!dst.empty()) R_CU(cudaMemcpyAsync(d_dst, dst.data(), dst.size() * 4, cudaMemcpyHostToDevice, r.stream));
  fill_links_kernel<<<(n + 255) / 256, 256, 0, r.stream>>>(r.adj, (uint32_t)h.M, d_start, d_dst, n);
  R_CU(cudaGetLastError());
  R_CU(cudaStreamSynchronize(r.stream));
  for (void* ptr : tmp) cudaFree(ptr);
  cudaSetDevice(prev);
  fnb_index_mutated(ix, true, false);
  return FNB_OK;
}

}  // extern "C"
