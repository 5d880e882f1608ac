// C++ surface of the B200 engine: a header-only shim with the shape of the reference's
// `flatnav::Index<dist_t, label_t>` (include/flatnav/index/Index.h:36-927 of BlaiseMuhirwa/flatnav), search side,
// written on top of the C ABI (include/flatnav_b200.h).  Code such as the reference's tools/query_npy.cpp:43-68
// or include/flatnav/tests/test_serialization.cpp:50-75 compiles against it after changing the include and the
// namespace (or with -DFLATNAV_B200_AS_FLATNAV, which aliases `flatnav` to this namespace):
//
//   auto index = Index<SquaredL2Distance<DataType::float32>, int>::loadIndex("graph.idx");
//   std::vector<std::pair<float, int>> top = index->search(query, /*K=*/10, /*ef_search=*/100);
//
// Kept from the reference: loadIndex / search / saveIndex / setNumThreads / getNumThreads and the getters
// (Index.h:442-531), their exceptions (std::runtime_error for I/O, std::invalid_argument for bad arguments)
// and move-only ownership; the constructing constructor with add / addBatch / allocateNode (Index.h:159-180,
// 262-378, GPU construction of csrc/build.cu), graph re-ordering and link import (Index.h:187-251, 412-440),
// distanceComputations / resetStats / getIndexSummary (Index.h:529-547).
// Added: searchBatch (what bindings.cpp:161-228 does with a thread pool).
#pragma once

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <iostream>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

#include "../flatnav_b200.h"

namespace flatnav_b200 {

namespace util {
// include/flatnav/util/Datatype.h:11-24 — the values are part of the file format
enum class DataType { uint8 = 0, uint16, uint32, uint64, int8, int16, int32, int64, float16, float32, float64, undefined };
}  // namespace util

namespace distances {
using util::DataType;
enum class MetricType { L2 = 0, IP = 1 };  // DistanceInterface.h

// Tag types standing where SquaredL2Distance<data_type> / InnerProductDistance<data_type>
// (distances/SquaredL2Distance.h:24, InnerProductDistance.h:23) stand in the reference: on the GPU the
// arithmetic lives in the kernels, the type only selects metric and element type.
// They keep what the reference's objects carry besides the arithmetic — the dimension — and its small interface
// (create / dimension / dataSize / getSummary, SquaredL2Distance.h:30-75), so construction code such as
// tools/construct_npy.cpp:78-80 (`auto distance = SquaredL2Distance<>::create(dim); Index<...>(std::move(distance),
// N, M)`) compiles unchanged.
inline size_t element_size(DataType t) { return t == DataType::float32 ? 4 : 1; }

template <DataType data_type = DataType::float32>
struct SquaredL2Distance {
  static constexpr int metric = FNB_METRIC_L2;
  static constexpr DataType dtype = data_type;
  SquaredL2Distance() = default;
  explicit SquaredL2Distance(size_t dim) : _dimension(dim) {}
  static std::unique_ptr<SquaredL2Distance<data_type>> create(size_t dim) {
    return std::make_unique<SquaredL2Distance<data_type>>(dim);
  }
  size_t dimension() const { return _dimension; }
  size_t dataSize() const { return _dimension * element_size(data_type); }
  void getSummary() const {
    std::cout << "\nSquaredL2Distance Parameters\n-----------------------------\n"
              << "Dimension: " << _dimension << "\n" << std::flush;
  }
  size_t _dimension = 0;
};
template <DataType data_type = DataType::float32>
struct InnerProductDistance {
  static constexpr int metric = FNB_METRIC_IP;
  static constexpr DataType dtype = data_type;
  InnerProductDistance() = default;
  explicit InnerProductDistance(size_t dim) : _dimension(dim) {}
  static std::unique_ptr<InnerProductDistance<data_type>> create(size_t dim) {
    return std::make_unique<InnerProductDistance<data_type>>(dim);
  }
  size_t dimension() const { return _dimension; }
  size_t dataSize() const { return _dimension * element_size(data_type); }
  void getSummary() const {
    std::cout << "\nInnerProductDistance Parameters\n-----------------------------\n"
              << "Dimension: " << _dimension << "\n" << std::flush;
  }
  size_t _dimension = 0;
};
}  // namespace distances

template <typename dist_t, typename label_t = int>
class Index {
  static_assert(std::is_same<label_t, int>::value || std::is_same<label_t, int32_t>::value,
                "the index file stores 32-bit labels (Index.h:566-573)");
  typedef std::pair<float, label_t> dist_label_t;
  typedef uint32_t node_id_t;

 public:
  // Index.h:159-180: an empty index for up to dataset_size vectors, built on the current CUDA device by add /
  // addBatch.  `dist` carries the dimension; `data_type` defaults to the distance's element type.
  Index(std::unique_ptr<dist_t> dist, int dataset_size, int max_edges_per_node, bool collect_stats = false,
        util::DataType data_type = dist_t::dtype)
      : _collect_stats(collect_stats) {
    if (!dist) throw std::invalid_argument("distance is null");
    raise(fnb_index_create(dist_t::metric, static_cast<int>(data_type), dist->dimension(),
                           static_cast<uint64_t>(std::max(dataset_size, 0)), static_cast<uint64_t>(std::max(max_edges_per_node, 0)),
                           -1, &_h));
    refresh();
  }

  Index(const Index&) = delete;
  Index& operator=(const Index&) = delete;
  Index(Index&& o) noexcept
      : _h(o._h), _info(o._info), _num_threads(o._num_threads), _collect_stats(o._collect_stats),
        _distance_computations(o._distance_computations.load()), _metric_hops(o._metric_hops.load()) {
    o._h = nullptr;
  }
  Index& operator=(Index&& o) noexcept {
    if (this != &o) {
      fnb_index_free(_h);
      _h = o._h;
      _info = o._info;
      _num_threads = o._num_threads;
      _collect_stats = o._collect_stats;
      _distance_computations = o._distance_computations.load();
      _metric_hops = o._metric_hops.load();
      o._h = nullptr;
    }
    return *this;
  }
  ~Index() { fnb_index_free(_h); }

  // Index.h:301-331 (+ add, selectNeighbors, connectNeighbors :341-378, 714-834): `data` is row-major
  // [labels.size(), dim] of data_type.  Insertion is batched on the GPU (csrc/build.cu); throws
  // std::invalid_argument for num_initializations <= 0 (:303-305) and std::runtime_error when the index is full
  // (:347-351).
  template <typename data_type>
  void addBatch(void* data, std::vector<label_t>& labels, int ef_construction, int num_initializations = 100) {
    static_assert(sizeof(label_t) == sizeof(int32_t), "32-bit labels");
    if (num_initializations <= 0) throw std::invalid_argument("num_initializations must be greater than 0.");
    if (sizeof(data_type) != distances::element_size(getDataType()))
      throw std::invalid_argument("data_type does not match the index data type");
    if (_info.cur_num_nodes + labels.size() > _info.max_node_count)
      throw std::runtime_error("Maximum number of nodes reached. Consider increasing the `max_node_count` parameter to "
                               "create a larger index.");
    raise(fnb_index_add(_h, data, reinterpret_cast<const int32_t*>(labels.data()), static_cast<int64_t>(labels.size()),
                        ef_construction, num_initializations, nullptr));
    refresh();
  }

  // Index.h:341-378: one vector.
  void add(void* data, label_t& label, int ef_construction, int num_initializations) {
    if (_info.cur_num_nodes >= _info.max_node_count)
      throw std::runtime_error("Maximum number of nodes reached. Consider increasing the `max_node_count` parameter to "
                               "create a larger index.");
    int32_t l = static_cast<int32_t>(label);
    raise(fnb_index_add(_h, data, &l, 1, ef_construction, num_initializations, nullptr));
    refresh();
  }

  // Index.h:262-272: appends an unlinked node (all link slots self-loops); out of room -> std::runtime_error.
  void allocateNode(void* data, label_t& label, node_id_t& new_node_id) {
    if (_info.cur_num_nodes >= _info.max_node_count)
      throw std::runtime_error("Maximum number of nodes reached. Consider increasing the `max_node_count` parameter to "
                               "create a larger index.");
    int32_t l = static_cast<int32_t>(label);
    new_node_id = static_cast<node_id_t>(_info.cur_num_nodes);
    raise(fnb_index_allocate_nodes(_h, data, &l, 1));
    refresh();
  }

  // Index.h:442-479.  `devices`: optional list of CUDA devices to replicate the index on.
  static std::unique_ptr<Index<dist_t, label_t>> loadIndex(const std::string& filename,
                                                           const std::vector<int>& devices = {}) {
    fnb_index* h = nullptr;
    int rc = fnb_index_load(filename.c_str(), dist_t::metric, static_cast<int>(dist_t::dtype),
                            devices.empty() ? nullptr : devices.data(), static_cast<int>(devices.size()), &h);
    raise(rc);
    std::unique_ptr<Index<dist_t, label_t>> index(new Index<dist_t, label_t>());
    index->_h = h;
    raise(fnb_index_info(h, &index->_info));
    index->_num_threads = std::max((uint32_t)1, (uint32_t)std::thread::hardware_concurrency() / 2);  // Index.h:467
    return index;
  }

  // Index.h:481-490
  void saveIndex(const std::string& filename) { raise(fnb_index_save(_h, filename.c_str())); }

  // Index.h:387-409: at most K (distance, label) pairs, ascending distance.
  std::vector<dist_label_t> search(const void* query, const int K, int ef_search, int num_initializations = 100) {
    std::vector<float> d(static_cast<size_t>(std::max(K, 0)));
    std::vector<int32_t> l(d.size());
    fnb_search_stats st{};
    int rc = fnb_search(_h, query, 1, K, ef_search, num_initializations, d.data(), l.data(), _collect_stats ? &st : nullptr);
    if (rc != FNB_SHORT_RESULT) raise(rc);
    count(st);
    std::vector<dist_label_t> out;
    out.reserve(d.size());
    for (size_t i = 0; i < d.size(); i++) {
      if (l[i] < 0 && d[i] == std::numeric_limits<float>::infinity()) break;  // fewer than K reachable
      out.emplace_back(d[i], static_cast<label_t>(l[i]));
    }
    return out;
  }

  // The batched fan-out of the Python binding (bindings.cpp:161-228) as one call: queries is row-major
  // [num_queries, dim] of the index element type; distances / labels are [num_queries, K].
  // Throws std::runtime_error if any query found fewer than K results (bindings.cpp:184-189).
  void searchBatch(const void* queries, size_t num_queries, int K, int ef_search, float* distances, label_t* labels,
                   int num_initializations = 100, fnb_search_stats* stats = nullptr) {
    fnb_search_stats st{};
    int rc = fnb_search(_h, queries, static_cast<int64_t>(num_queries), K, ef_search, num_initializations, distances,
                        reinterpret_cast<int32_t*>(labels), &st);
    if (stats) *stats = st;
    if (rc == FNB_OK || rc == FNB_SHORT_RESULT) count(st);
    raise(rc);
  }

  // Index.h:412-440 — the ordering runs on the host (same queue discipline as util::gOrder / util::rcmOrder, so the
  // permutation is the reference's), the relabelling and re-layout on the GPU (csrc/reorder.cu).
  void doGraphReordering(const std::vector<std::string>& reordering_methods) {
    for (const auto& method : reordering_methods) {
      if (method == "gorder") {
        reorderGOrder(5);
      } else if (method == "rcm") {
        reorderRCM();
      } else {
        throw std::invalid_argument("Invalid reordering method: " + method);
      }
    }
  }
  void reorderGOrder(const int window_size = 5) { raise(fnb_index_reorder(_h, FNB_REORDER_GORDER, window_size, nullptr)); }
  void reorderRCM() { raise(fnb_index_reorder(_h, FNB_REORDER_RCM, 0, nullptr)); }

  // Index.h:240-251
  std::vector<std::vector<uint32_t>> getGraphOutdegreeTable() {
    const size_t n = _info.cur_num_nodes, M = _info.max_edges_per_node;
    std::vector<uint32_t> links(n * M);
    raise(fnb_index_links(_h, links.data()));
    std::vector<std::vector<uint32_t>> table(n);
    for (size_t node = 0; node < n; node++)
      for (size_t i = 0; i < M; i++)
        if (links[node * M + i] != node) table[node].push_back(links[node * M + i]);
    return table;
  }

  // Index.h:187-238
  void buildGraphLinks(const std::string& mtx_filename) { raise(fnb_index_build_graph_links(_h, mtx_filename.c_str())); }

  // Index.h:492-503 — validated like the reference; GPU execution does not use it.
  inline void setNumThreads(uint32_t num_threads) {
    if (num_threads == 0 || num_threads > std::thread::hardware_concurrency()) {
      throw std::invalid_argument(
          "Number of threads must be greater than 0 and less than or equal to the number of hardware threads.");
    }
    _num_threads = num_threads;
  }
  inline uint32_t getNumThreads() const { return _num_threads; }

  // Index.h:519-531
  inline size_t maxEdgesPerNode() const { return _info.max_edges_per_node; }
  inline size_t dataSizeBytes() const { return _info.data_size_bytes; }
  inline size_t nodeSizeBytes() const { return _info.node_size_bytes; }
  inline size_t maxNodeCount() const { return _info.max_node_count; }
  inline size_t currentNumNodes() const { return _info.cur_num_nodes; }
  inline size_t dataDimension() const { return _info.dim; }
  inline util::DataType getDataType() const { return static_cast<util::DataType>(_info.data_type); }
  inline uint64_t getTotalIndexMemory() const { return _info.node_size_bytes * _info.max_node_count; }
  inline const fnb_info& deviceInfo() const { return _info; }
  fnb_index* handle() { return _h; }

  // Index.h:529, 533-536: counted only with collect_stats, like the reference; the numbers are the kernel's own
  // counters (every database-row evaluation incl. the entry probes / every expanded node).
  inline uint64_t distanceComputations() const { return _distance_computations; }
  inline uint64_t metricHops() const { return _metric_hops; }
  void resetStats() {
    _distance_computations = 0;
    _metric_hops = 0;
  }

  // Index.h:538-547
  void getIndexSummary() const {
    std::cout << "\nIndex Parameters\n" << std::flush;
    std::cout << "-----------------------------\n" << std::flush;
    std::cout << "max_edges_per_node (M): " << _info.max_edges_per_node << "\n" << std::flush;
    std::cout << "data_size_bytes: " << _info.data_size_bytes << "\n" << std::flush;
    std::cout << "node_size_bytes: " << _info.node_size_bytes << "\n" << std::flush;
    std::cout << "max_node_count: " << _info.max_node_count << "\n" << std::flush;
    std::cout << "cur_num_nodes: " << _info.cur_num_nodes << "\n" << std::flush;
    dist_t(_info.dim).getSummary();
  }

 private:
  Index() = default;
  void refresh() { raise(fnb_index_info(_h, &_info)); }
  void count(const fnb_search_stats& st) {
    if (!_collect_stats) return;
    _distance_computations += static_cast<uint64_t>(st.n_dist);
    _metric_hops += static_cast<uint64_t>(st.n_hops);
  }
  static void raise(int rc) {
    if (rc == FNB_OK) return;
    const std::string msg = fnb_last_error();
    if (rc == FNB_ERR_INVALID_ARG) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
  }
  fnb_index* _h = nullptr;
  fnb_info _info{};
  uint32_t _num_threads = 1;
  bool _collect_stats = false;
  // atomics, as in the reference: search() is re-entrant and callers fan it out over threads
  std::atomic<uint64_t> _distance_computations{0};
  std::atomic<uint64_t> _metric_hops{0};
};

}  // namespace flatnav_b200

#ifdef FLATNAV_B200_AS_FLATNAV
namespace flatnav = flatnav_b200;
#endif
