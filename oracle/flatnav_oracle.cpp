// TEST INFRASTRUCTURE — CPU restatement ("oracle") of FlatNav's search hot path.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library, and only as the checker / baseline.  The product (flatnav_b200) never links,
// imports or falls back to anything in oracle/.
//
// Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this restatement against the
// unmodified reference compiled from /root/reference (oracle/_ref/ref_flatnav*, see Makefile) and
// against golden outputs of that binary committed under tests/golden/ (made by
// tools/make_golden.py).
//
// What is restated (all paths relative to /root/reference):
//   ora_search mode 0 ("heaps")  — Index::search            include/flatnav/index/Index.h:387-409
//                                  Index::initializeSearch  Index.h:845-870
//                                  Index::beamSearch        Index.h:606-659
//                                  Index::processCandidateNode Index.h:661-707
//                                  PriorityQueue/CompareByFirst Index.h:47-53
//                                  VisitedSet semantics     include/flatnav/util/VisitedSetPool.h:36-48
//                                  node accessors           Index.h:555-573
//   ora_search mode 1 ("list")   — the single sorted-list formulation the CUDA kernel implements
//                                  (SURVEY.md §8c): same visit order and acceptance rule, batch merge
//                                  of one expansion's fresh neighbours, ties ordered by (distance, id).
//                                  Identical to mode 0 except on exact distance ties.
//   distances                    — defaultSquaredL2 / defaultInnerProduct
//                                  include/flatnav/distances/L2DistanceDispatcher.h:9-17
//                                  include/flatnav/distances/IPDistanceDispatcher.h:9-16
//                                  (uint8/int8: integer products accumulated, result cast to float;
//                                   float32: see `dist_order` below)
//   ora_parse_header             — Index::serialize / loadIndex byte layout  Index.h:134-141, 442-479
//   ora_bruteforce               — exact scan (not in the reference; parity unpinned BY THE
//                                  REFERENCE for this one function — pinned by a numpy float64 scan
//                                  in tests instead), ties → lower node id.
//
//   ora_reorder                  — Index::doGraphReordering one step (gorder window w / rcm)
//                                  Index.h:412-440; util::gOrder, util::rcmOrder Reordering.h:26-199;
//                                  GorderPriorityQueue.h:13-112 (sorted vector + std::upper_bound /
//                                  lower_bound, restated literally; the product uses an O(1) boundary
//                                  table instead); Index::relabel Index.h:872-926 applied to the file
//                                  image in place.  Pinned to the reference: byte-identical files.
//   ora_build_graph_links        — Index::buildGraphLinks edge rule Index.h:219-234 (edges given as arrays).
//
// dist_order (float32 only; integer types are exact in any order):
//   0 = "sequential": the scalar definition, one float accumulator, elements in index order,
//       separate multiply and add (what defaultSquaredL2/defaultInnerProduct spell out).
//   1 = "lanes": the summation order of the CUDA kernel (flatnav_b200/csrc/fnb_layout.h): the row
//       is cut into 16-byte chunks; lane p of G accumulates chunks p, p+G, p+2G, ... with fused
//       multiply-add, element by element; lanes are then combined by an xor butterfly with offsets
//       G/2, G/4, ..., 1.  G = 8 when the row has <= 32 chunks, else 32.
//   The reference's own AVX-512/AVX/SSE kernels use yet another association (16 lanes, tree
//   reduce, -ffast-math); all three agree to ~1e-6 relative, the parity tolerance is 1e-5.

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <queue>
#include <thread>
#include <utility>
#include <vector>

extern "C" {

// flatnav::util::DataType values that can appear in an index file (util/Datatype.h:11-24)
enum { ORA_U8 = 0, ORA_I8 = 4, ORA_F32 = 9 };
enum { ORA_L2 = 0, ORA_IP = 1 };

struct ora_index {
  int32_t data_type;  // ORA_U8 / ORA_I8 / ORA_F32
  int32_t metric;     // ORA_L2 / ORA_IP (NOT stored in the file; implied by the C++ type in the reference)
  uint64_t M;
  uint64_t data_size_bytes;
  uint64_t node_size_bytes;
  uint64_t max_node_count;
  uint64_t cur_num_nodes;
  uint64_t dim;
  const uint8_t* mem;  // node blob: [vector | M x u32 links | i32 label] * max_node_count
};

// Header = int32 data_type, u64 M, u64 data_size, u64 node_size, u64 max_nodes, u64 cur_nodes,
// then the distance object: u64 dimension, u64 data_size  => 60 bytes, little endian, no padding.
int ora_parse_header(const void* file_bytes, uint64_t nbytes, int metric, ora_index* out) {
  if (nbytes < 60) return -1;
  const uint8_t* p = static_cast<const uint8_t*>(file_bytes);
  int32_t dt;
  uint64_t v[7];
  std::memcpy(&dt, p, 4);
  std::memcpy(v, p + 4, 56);
  out->data_type = dt;
  out->metric = metric;
  out->M = v[0];
  out->data_size_bytes = v[1];
  out->node_size_bytes = v[2];
  out->max_node_count = v[3];
  out->cur_num_nodes = v[4];
  out->dim = v[5];
  if (v[6] != v[1]) return -2;
  if (out->node_size_bytes != out->data_size_bytes + 4 * out->M + 4) return -3;
  uint64_t es = dt == ORA_F32 ? 4 : (dt == ORA_U8 || dt == ORA_I8) ? 1 : 0;
  if (es == 0 || out->dim * es != out->data_size_bytes) return -4;
  if (nbytes < 60 + out->node_size_bytes * out->max_node_count) return -5;
  out->mem = p + 60;
  return 0;
}

}  // extern "C"

namespace {

inline const uint8_t* node_data(const ora_index& ix, uint32_t n) { return ix.mem + (uint64_t)n * ix.node_size_bytes; }
inline const uint32_t* node_links(const ora_index& ix, uint32_t n) {
  return reinterpret_cast<const uint32_t*>(node_data(ix, n) + ix.data_size_bytes);
}
inline int32_t node_label(const ora_index& ix, uint32_t n) {
  int32_t l;
  std::memcpy(&l, node_data(ix, n) + ix.data_size_bytes + 4 * ix.M, 4);
  return l;
}

// ---- float32 distances ---------------------------------------------------------------------
template <bool IP>
float dist_f32_sequential(const float* x, const float* y, size_t d) {
  volatile float acc = 0.f;  // volatile: forbid re-association / vectorisation of the definition
  for (size_t i = 0; i < d; i++) {
    if (IP) {
      float prod = x[i] * y[i];
      acc = acc + prod;
    } else {
      float diff = x[i] - y[i];
      float sq = diff * diff;
      acc = acc + sq;
    }
  }
  return IP ? 1.0f - acc : (float)acc;
}

template <bool IP>
float dist_f32_lanes(const float* x, const float* y, size_t d) {
  size_t nchunks = (d + 3) / 4;
  int G = nchunks <= 32 ? 8 : 32;
  float lane[32];
  for (int p = 0; p < G; p++) {
    float acc = 0.f;
    for (size_t c = (size_t)p; c < nchunks; c += (size_t)G) {
      for (size_t e = 4 * c; e < 4 * c + 4 && e < d; e++) {
        if (IP) {
          acc = std::fmaf(x[e], y[e], acc);
        } else {
          float diff = x[e] - y[e];
          acc = std::fmaf(diff, diff, acc);
        }
      }
    }
    lane[p] = acc;
  }
  for (int off = G / 2; off > 0; off >>= 1) {
    float nxt[32];
    for (int p = 0; p < G; p++) nxt[p] = lane[p] + lane[p ^ off];
    std::memcpy(lane, nxt, sizeof(float) * (size_t)G);
  }
  return IP ? 1.0f - lane[0] : lane[0];
}

// ---- uint8 / int8 distances: integer arithmetic per element, float result -------------------
template <typename T, bool IP>
float dist_int(const T* x, const T* y, size_t d) {
  // The reference accumulates integer-valued terms in a float (exact while the running sum stays
  // below 2^24, i.e. for every BASELINE config); an int64 accumulator states the same value
  // without the rounding caveat.
  int64_t acc = 0;
  for (size_t i = 0; i < d; i++) {
    int a = (int)x[i], b = (int)y[i];
    acc += IP ? (int64_t)(a * b) : (int64_t)((a - b) * (a - b));
  }
  return IP ? 1.0f - (float)acc : (float)acc;
}

struct Dist {
  const ora_index& ix;
  int order;
  float operator()(const void* q, const void* row) const {
    size_t d = ix.dim;
    bool ip = ix.metric == ORA_IP;
    switch (ix.data_type) {
      case ORA_F32: {
        const float* a = static_cast<const float*>(q);
        const float* b = static_cast<const float*>(row);
        if (order == 0) return ip ? dist_f32_sequential<true>(a, b, d) : dist_f32_sequential<false>(a, b, d);
        return ip ? dist_f32_lanes<true>(a, b, d) : dist_f32_lanes<false>(a, b, d);
      }
      case ORA_U8:
        return ip ? dist_int<uint8_t, true>((const uint8_t*)q, (const uint8_t*)row, d)
                  : dist_int<uint8_t, false>((const uint8_t*)q, (const uint8_t*)row, d);
      default:
        return ip ? dist_int<int8_t, true>((const int8_t*)q, (const int8_t*)row, d)
                  : dist_int<int8_t, false>((const int8_t*)q, (const int8_t*)row, d);
    }
  }
};

struct Counters {
  int64_t ndist = 0;  // database-row distance evaluations, INCLUDING the entry-selection probes
  int64_t nhops = 0;  // expanded nodes
};

// Index::initializeSearch (Index.h:845-870): strided probes, first strict minimum wins.
uint32_t initialize_search(const ora_index& ix, const Dist& dist, const void* q, int ninit, Counters& c) {
  int step = (int)(ix.cur_num_nodes / (uint64_t)ninit);
  step = step ? step : 1;
  float min_dist = std::numeric_limits<float>::max();
  uint32_t entry = 0;
  for (uint64_t node = 0; node < ix.cur_num_nodes; node += (uint64_t)step) {
    float d = dist(q, node_data(ix, (uint32_t)node));
    c.ndist++;
    if (d < min_dist) {
      min_dist = d;
      entry = (uint32_t)node;
    }
  }
  return entry;
}

typedef std::pair<float, uint32_t> dist_node_t;
struct CompareByFirst {
  bool operator()(dist_node_t const& a, dist_node_t const& b) const noexcept { return a.first < b.first; }
};
typedef std::priority_queue<dist_node_t, std::vector<dist_node_t>, CompareByFirst> PriorityQueue;

// mode 0: the reference's two-heap loop, statement for statement (Index.h:606-707, 387-409).
void search_heaps(const ora_index& ix, const Dist& dist, const void* q, int K, int ef, int ninit,
                  std::vector<uint8_t>& visited, std::vector<std::pair<float, int32_t>>& out, Counters& c) {
  uint32_t entry = initialize_search(ix, dist, q, ninit, c);
  size_t buffer_size = (size_t)std::max(ef, K);
  PriorityQueue neighbors, candidates;
  std::fill(visited.begin(), visited.end(), 0);  // VisitedSet::clear()

  float d0 = dist(q, node_data(ix, entry));
  // (this re-evaluation of the entry node is not counted: the reference counts only
  //  processCandidateNode evaluations plus `ninit`, Index.h:689-691, 857-859)
  float max_dist = d0;
  candidates.emplace(-d0, entry);
  neighbors.emplace(d0, entry);
  visited[entry] = 1;

  while (!candidates.empty()) {
    auto top = candidates.top();
    if (-top.first > max_dist && neighbors.size() >= buffer_size) break;
    candidates.pop();
    c.nhops++;
    const uint32_t* links = node_links(ix, top.second);
    for (uint64_t i = 0; i < ix.M; i++) {
      uint32_t nb = links[i];
      if (visited[nb]) continue;
      visited[nb] = 1;
      float d = dist(q, node_data(ix, nb));
      c.ndist++;
      if (neighbors.size() < buffer_size || d < max_dist) {
        candidates.emplace(-d, nb);
        neighbors.emplace(d, nb);
        if (neighbors.size() > buffer_size) neighbors.pop();
        if (!neighbors.empty()) max_dist = neighbors.top().first;
      }
    }
  }
  out.clear();
  while (!neighbors.empty()) {
    out.emplace_back(neighbors.top().first, node_label(ix, neighbors.top().second));
    neighbors.pop();
  }
  std::sort(out.begin(), out.end(),
            [](const std::pair<float, int32_t>& l, const std::pair<float, int32_t>& r) { return l.first < r.first; });
  if (out.size() > (size_t)K) out.resize((size_t)K);
}

// mode 1: sorted-list formulation (what the CUDA kernel does).
struct Entry {
  float d;
  uint32_t id;
  bool expanded;
};
inline bool key_less(const Entry& a, const Entry& b) { return a.d < b.d || (a.d == b.d && a.id < b.id); }

void search_list(const ora_index& ix, const Dist& dist, const void* q, int K, int ef, int ninit,
                 std::vector<uint8_t>& visited, std::vector<std::pair<float, int32_t>>& out, Counters& c) {
  uint32_t entry = initialize_search(ix, dist, q, ninit, c);
  size_t B = (size_t)std::max(ef, K);
  std::fill(visited.begin(), visited.end(), 0);
  std::vector<Entry> L;
  L.reserve(B + ix.M);
  L.push_back({dist(q, node_data(ix, entry)), entry, false});
  visited[entry] = 1;
  std::vector<Entry> fresh;
  for (;;) {
    size_t pick = L.size();
    for (size_t i = 0; i < L.size(); i++)
      if (!L[i].expanded) {
        pick = i;
        break;
      }
    if (pick == L.size()) break;  // every entry of the list is expanded
    L[pick].expanded = true;
    c.nhops++;
    const uint32_t* links = node_links(ix, L[pick].id);
    // Links are taken in groups of 32 (one warp-wide load in the CUDA kernel): acceptance is tested against
    // the list as it stood before the group, then the group's accepted entries are merged at once.  For
    // M <= 32 this is one batch per expansion; it differs from link-by-link processing only on exact ties.
    for (uint64_t l0 = 0; l0 < ix.M; l0 += 32) {
      bool full = L.size() >= B;
      float worst = L.back().d;
      fresh.clear();
      for (uint64_t i = l0; i < ix.M && i < l0 + 32; i++) {
        uint32_t nb = links[i];
        if (visited[nb]) continue;
        visited[nb] = 1;
        float d = dist(q, node_data(ix, nb));
        c.ndist++;
        if (!full || d < worst) fresh.push_back({d, nb, false});
      }
      if (fresh.empty()) continue;
      std::sort(fresh.begin(), fresh.end(), key_less);
      std::vector<Entry> merged(L.size() + fresh.size());
      std::merge(L.begin(), L.end(), fresh.begin(), fresh.end(), merged.begin(), key_less);
      if (merged.size() > B) merged.resize(B);
      L.swap(merged);
    }
  }
  out.clear();
  for (size_t i = 0; i < L.size() && i < (size_t)K; i++) out.emplace_back(L[i].d, node_label(ix, L[i].id));
}

}  // namespace

extern "C" {

// Batched search.  out_dist/out_label are [Q,K]; unfilled slots get +inf / -1.
// out_ndist/out_nhops (nullable) are per-query counters.  Returns number of short results, <0 on error.
int64_t ora_search(const ora_index* ixp, const void* queries, int64_t Q, int K, int ef, int ninit, int mode,
                   int dist_order, int threads, float* out_dist, int32_t* out_label, int64_t* out_ndist,
                   int64_t* out_nhops) {
  if (ninit <= 0) return -10;  // std::invalid_argument in the reference (Index.h:847-849)
  const ora_index& ix = *ixp;
  Dist dist{ix, dist_order};
  std::atomic<int64_t> next(0), shorts(0);
  if (threads < 1) threads = 1;
  auto worker = [&]() {
    std::vector<uint8_t> visited(ix.max_node_count);
    std::vector<std::pair<float, int32_t>> res;
    for (;;) {
      int64_t i = next.fetch_add(1);
      if (i >= Q) break;
      const uint8_t* q = static_cast<const uint8_t*>(queries) + (uint64_t)i * ix.data_size_bytes;
      Counters c;
      if (mode == 0)
        search_heaps(ix, dist, q, K, ef, ninit, visited, res, c);
      else
        search_list(ix, dist, q, K, ef, ninit, visited, res, c);
      for (int j = 0; j < K; j++) {
        bool have = (size_t)j < res.size();
        out_dist[i * K + j] = have ? res[(size_t)j].first : std::numeric_limits<float>::infinity();
        out_label[i * K + j] = have ? res[(size_t)j].second : -1;
      }
      if (res.size() < (size_t)K) shorts.fetch_add(1);
      if (out_ndist) out_ndist[i] = c.ndist;
      if (out_nhops) out_nhops[i] = c.nhops;
    }
  };
  if (threads == 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
  }
  return shorts.load();
}

// Single distance evaluation (for unit tests of the distance definitions).
float ora_distance(int data_type, int metric, int dist_order, const void* x, const void* y, uint64_t dim) {
  ora_index ix{};
  ix.data_type = data_type;
  ix.metric = metric;
  ix.dim = dim;
  return Dist{ix, dist_order}(x, y);
}

// Exact scan over nodes [0, cur_num_nodes): top-K by (distance, node id); labels from the label field.
int64_t ora_bruteforce(const ora_index* ixp, const void* queries, int64_t Q, int K, int dist_order, int threads,
                       float* out_dist, int32_t* out_label) {
  const ora_index& ix = *ixp;
  Dist dist{ix, dist_order};
  std::atomic<int64_t> next(0);
  if (threads < 1) threads = 1;
  auto worker = [&]() {
    std::vector<std::pair<float, uint32_t>> all(ix.cur_num_nodes);
    for (;;) {
      int64_t i = next.fetch_add(1);
      if (i >= Q) break;
      const uint8_t* q = static_cast<const uint8_t*>(queries) + (uint64_t)i * ix.data_size_bytes;
      for (uint64_t n = 0; n < ix.cur_num_nodes; n++) all[n] = {dist(q, node_data(ix, (uint32_t)n)), (uint32_t)n};
      size_t k = std::min<size_t>((size_t)K, all.size());
      std::partial_sort(all.begin(), all.begin() + (long)k, all.end());
      for (int j = 0; j < K; j++) {
        bool have = (size_t)j < k;
        out_dist[i * K + j] = have ? all[(size_t)j].first : std::numeric_limits<float>::infinity();
        out_label[i * K + j] = have ? node_label(ix, all[(size_t)j].second) : -1;
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) pool.emplace_back(worker);
  for (auto& t : pool) t.join();
  return 0;
}

}  // extern "C"

// =================================================================================================
// Graph re-ordering (Index.h:412-440, util/Reordering.h, util/GorderPriorityQueue.h) on a file image
// =================================================================================================
namespace {

typedef std::vector<std::vector<uint32_t>> table_t;

// Index::getGraphOutdegreeTable (Index.h:240-251)
table_t outdegree_table(const ora_index& ix) {
  table_t t(ix.cur_num_nodes);
  for (uint32_t n = 0; n < ix.cur_num_nodes; n++) {
    const uint32_t* links = node_links(ix, n);
    for (uint64_t i = 0; i < ix.M; i++)
      if (links[i] != n) t[n].push_back(links[i]);
  }
  return t;
}

// GorderPriorityQueue.h:13-112: entries sorted ascending by priority; the reference's hash map
// key -> slot is a plain vector here (-1 = erased), everything else as written there.
struct GorderPQ {
  struct Node {
    uint32_t key;
    int priority;
  };
  std::vector<Node> list;
  std::vector<int64_t> index;
  static bool compare(const Node& a, const Node& b) { return a.priority < b.priority; }
  explicit GorderPQ(size_t n) : list(n), index(n) {
    for (size_t i = 0; i < n; i++) {
      list[i] = Node{(uint32_t)i, 0};
      index[i] = (int64_t)i;
    }
  }
  void swap(size_t i, size_t j) {
    Node tmp = list[i];
    list[i] = list[j];
    list[j] = tmp;
    index[list[i].key] = (int64_t)i;
    index[list[j].key] = (int64_t)j;
  }
  void increment(uint32_t key) {
    if (index[key] < 0) return;
    const size_t i = (size_t)index[key];
    auto it = std::upper_bound(list.begin(), list.end(), list[i], compare);
    const size_t new_index = (size_t)(it - list.begin()) - 1;
    swap(i, new_index);
    list[new_index].priority++;
  }
  void decrement(uint32_t key) {
    if (index[key] < 0) return;
    const size_t i = (size_t)index[key];
    auto it = std::lower_bound(list.begin(), list.end(), list[i], compare);
    const size_t new_index = (size_t)(it - list.begin());
    swap(i, new_index);
    list[new_index].priority--;
  }
  uint32_t pop() {
    Node max = list.back();
    list.pop_back();
    index[max.key] = -1;
    return max.key;
  }
};

// util::gOrder (Reordering.h:26-117)
std::vector<uint32_t> gorder(const table_t& out, int w) {
  const int64_t n = (int64_t)out.size();
  table_t in(n);
  for (int64_t node = 0; node < n; node++)
    for (uint32_t edge : out[node]) in[edge].push_back((uint32_t)node);
  GorderPQ Q(n);
  std::vector<uint32_t> P(n, 0);
  if (n == 0) return P;
  Q.increment(0);
  P[0] = Q.pop();
  for (int64_t i = 1; i < n; i++) {
    const uint32_t v_e = P[i - 1];
    for (uint32_t u : out[v_e]) Q.increment(u);
    for (uint32_t u : in[v_e]) {
      Q.increment(u);
      for (uint32_t v : out[u]) Q.increment(v);
    }
    if (i > w + 1) {
      const uint32_t v_b = P[i - w - 1];
      for (uint32_t u : out[v_b]) Q.decrement(u);
      for (uint32_t u : in[v_b]) {
        Q.decrement(u);
        for (uint32_t v : out[u]) Q.decrement(v);
      }
    }
    P[i] = Q.pop();
  }
  std::vector<uint32_t> Pinv(n, 0);
  for (int64_t k = 0; k < n; k++) Pinv[P[k]] = (uint32_t)k;
  return Pinv;
}

// util::rcmOrder (Reordering.h:119-199)
std::vector<uint32_t> rcm(const table_t& out) {
  const size_t n = out.size();
  typedef std::pair<uint32_t, int> nd_t;
  auto less_degree = [](const nd_t& a, const nd_t& b) { return a.second < b.second; };
  std::vector<nd_t> sorted_nodes;
  std::vector<int> degrees;
  for (size_t node = 0; node < n; node++) {
    const int deg = (int)out[node].size();
    sorted_nodes.push_back({(uint32_t)node, deg});
    degrees.push_back(deg);
  }
  std::sort(sorted_nodes.begin(), sorted_nodes.end(), less_degree);
  std::vector<uint32_t> P;
  std::vector<bool> visited(n, false);
  for (size_t i = 0; i < sorted_nodes.size(); i++) {
    const uint32_t node = sorted_nodes[i].first;
    std::queue<uint32_t> Q;
    if (visited[node]) continue;
    P.push_back(node);
    visited[node] = true;
    std::vector<nd_t> neighbors;
    for (uint32_t edge : out[node]) neighbors.push_back({edge, degrees[edge]});
    std::sort(neighbors.begin(), neighbors.end(), less_degree);
    for (auto& x : neighbors) Q.push(x.first);
    while (!Q.empty()) {
      const uint32_t candidate = Q.front();
      Q.pop();
      if (visited[candidate]) continue;
      P.push_back(candidate);
      visited[candidate] = true;
      std::vector<nd_t> cn;
      for (uint32_t edge : out[candidate]) cn.push_back({edge, degrees[edge]});
      std::sort(cn.begin(), cn.end(), less_degree);
      for (auto& x : cn) Q.push(x.first);
    }
  }
  std::reverse(P.begin(), P.end());
  std::vector<uint32_t> Pinv(n, 0);
  for (size_t k = 0; k < n; k++) Pinv[P[k]] = (uint32_t)k;
  return Pinv;
}

// Index::relabel (Index.h:872-926) + swapNodes (:575-592) on the node blob, in place
void relabel(const ora_index& ix, uint8_t* mem, const std::vector<uint32_t>& P) {
  const uint64_t ns = ix.node_size_bytes;
  for (uint32_t n = 0; n < ix.cur_num_nodes; n++) {
    uint32_t* links = reinterpret_cast<uint32_t*>(mem + (uint64_t)n * ns + ix.data_size_bytes);
    for (uint64_t m = 0; m < ix.M; m++) links[m] = P[links[m]];
  }
  std::vector<uint8_t> temp(ns);
  auto swap_nodes = [&](uint32_t a, uint32_t b) {  // whole node = data + links + label
    std::memcpy(temp.data(), mem + (uint64_t)b * ns, ns);
    std::memmove(mem + (uint64_t)b * ns, mem + (uint64_t)a * ns, ns);
    std::memcpy(mem + (uint64_t)a * ns, temp.data(), ns);
  };
  std::vector<bool> relocated(ix.cur_num_nodes, false);
  for (uint32_t n = 0; n < ix.cur_num_nodes; n++) {
    if (relocated[n]) continue;
    const uint32_t src = n;
    uint32_t dest = P[src];
    swap_nodes(src, dest);
    relocated[src] = true;
    while (!relocated[dest]) {
      relocated[dest] = true;
      dest = P[dest];
      swap_nodes(src, dest);
    }
  }
}

}  // namespace

extern "C" {

// One step of Index::doGraphReordering on a WRITABLE file image.  method 0 = gorder (window w), 1 = rcm.
// perm_out (nullable): uint32 [cur_num_nodes], old id -> new id.
int ora_reorder(void* file_bytes, uint64_t nbytes, int method, int w, uint32_t* perm_out) {
  ora_index ix;
  int rc = ora_parse_header(file_bytes, nbytes, ORA_L2, &ix);
  if (rc != 0) return rc;
  if (method != 0 && method != 1) return -20;
  table_t out = outdegree_table(ix);
  std::vector<uint32_t> P = method == 0 ? gorder(out, w) : rcm(out);
  relabel(ix, static_cast<uint8_t*>(file_bytes) + 60, P);
  if (perm_out)
    for (size_t i = 0; i < P.size(); i++) perm_out[i] = P[i];
  return 0;
}

// Edge rule of Index::buildGraphLinks (Index.h:219-234) on a WRITABLE file image; edges 0-based, file order.
int ora_build_graph_links(void* file_bytes, uint64_t nbytes, const uint32_t* src, const uint32_t* dst, uint64_t n_edges) {
  ora_index ix;
  int rc = ora_parse_header(file_bytes, nbytes, ORA_L2, &ix);
  if (rc != 0) return rc;
  uint8_t* mem = static_cast<uint8_t*>(file_bytes) + 60;
  for (uint64_t e = 0; e < n_edges; e++) {
    const uint32_t u = src[e], v = dst[e];
    uint32_t* links = reinterpret_cast<uint32_t*>(mem + (uint64_t)u * ix.node_size_bytes + ix.data_size_bytes);
    for (uint64_t i = 0; i < ix.M; i++) {
      if (links[i] == u) {
        links[i] = v;
        break;
      }
    }
  }
  return 0;
}

}  // extern "C"
