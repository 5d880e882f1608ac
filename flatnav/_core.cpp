// `flatnav._core` — the reference's Python extension module (python-bindings/src/flatnav/bindings.cpp:426-538) rebuilt
// as a thin pybind11 layer over the C ABI of libflatnav_b200.so (include/flatnav_b200.h).  Same module layout
// (`_core.index`, `_core.data_type`, `_core.MetricType`, `__version__`), same six index classes, method names, keyword
// names, defaults, return dtypes / shapes and exception types, so that `import flatnav` code written against the
// reference runs on the B200 engine unchanged.  Nothing is computed here: every method forwards to an fnb_* entry
// point; without the shared library or a CUDA device the calls fail loudly (there is no CPU path).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cstdint>
#include <iostream>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../include/flatnav_b200.h"

namespace py = pybind11;

namespace {

// flatnav::util::DataType values the binding exports (util/Datatype.h:11-24)
enum class DataType : int { uint8 = FNB_DTYPE_UINT8, int8 = FNB_DTYPE_INT8, float32 = FNB_DTYPE_FLOAT32 };
enum class MetricType : int { L2 = FNB_METRIC_L2, IP = FNB_METRIC_IP };

void check(int rc) {
  if (rc == FNB_OK) return;
  const std::string msg = fnb_last_error();
  if (rc == FNB_ERR_INVALID_ARG) throw std::invalid_argument(msg);  // -> ValueError
  if (rc == FNB_ERR_NOMEM) throw std::bad_alloc();
  throw std::runtime_error(msg);  // short results included (bindings.cpp:134-137, 184-189)
}

size_t elem_size(DataType dt) { return dt == DataType::float32 ? 4 : 1; }

// py::array_t<T, c_style | forcecast> of the index's element type (bindings.cpp:36-51)
py::array as_index_dtype(const py::array& a, DataType dt) {
  switch (dt) {
    case DataType::float32: return a.cast<py::array_t<float, py::array::c_style | py::array::forcecast>>();
    case DataType::int8: return a.cast<py::array_t<int8_t, py::array::c_style | py::array::forcecast>>();
    case DataType::uint8: return a.cast<py::array_t<uint8_t, py::array::c_style | py::array::forcecast>>();
  }
  throw std::invalid_argument("Unsupported data type.");
}

class GpuIndex : public std::enable_shared_from_this<GpuIndex> {
 public:
  GpuIndex(fnb_index* h, MetricType metric, DataType dt, bool verbose) : _h(h), _metric(metric), _dt(dt), _verbose(verbose) {
    fnb_info info;
    check(fnb_index_info(_h, &info));
    _dim = (int)info.dim;
    _num_threads = std::max(1u, std::thread::hardware_concurrency() / 2);  // Index.h:467
    if (_verbose) summary();
  }
  GpuIndex(const GpuIndex&) = delete;
  GpuIndex& operator=(const GpuIndex&) = delete;
  virtual ~GpuIndex() { fnb_index_free(_h); }

  void summary() const {  // Index::getIndexSummary (Index.h:538-547)
    fnb_info info;
    check(fnb_index_info(_h, &info));
    std::cout << "\nIndex Parameters\n-----------------------------\n"
              << "max_edges_per_node (M): " << info.max_edges_per_node << "\n"
              << "data_size_bytes: " << info.data_size_bytes << "\n"
              << "node_size_bytes: " << info.node_size_bytes << "\n"
              << "max_node_count: " << info.max_node_count << "\n"
              << "cur_num_nodes: " << info.cur_num_nodes << "\n"
              << "dimension: " << info.dim << "\n"
              << "engine: " << fnb_version() << "\n"
              << std::flush;
  }

  // PyIndex::add (bindings.cpp:62-110, 326-335)
  void add(const py::array& data, int ef_construction, int num_initializations, py::object labels) {
    py::array a = as_index_dtype(data, _dt);
    if (a.ndim() != 2 || a.shape(1) != _dim)
      throw std::invalid_argument("Data has incorrect dimensions. data.ndim() = `" + std::to_string(a.ndim()) +
                                  "` and data_dim = `" + std::to_string(a.ndim() > 1 ? a.shape(1) : -1) +
                                  "`. Expected 2D array with dimensions (num_vectors, dim).");
    const int64_t n = a.shape(0);
    std::vector<int32_t> lab;
    if (!labels.is_none()) {
      try {
        lab = py::cast<std::vector<int32_t>>(labels);
      } catch (const py::cast_error&) {
        throw std::invalid_argument("Invalid labels provided.");
      }
      if ((int64_t)lab.size() != n) throw std::invalid_argument("Incorrect number of labels.");
    }
    int rc;
    {
      py::gil_scoped_release gil;
      rc = fnb_index_add(_h, a.data(), lab.empty() ? nullptr : lab.data(), n, ef_construction, num_initializations, nullptr);
    }
    check(rc);
  }

  // PyIndex::allocateNodes (bindings.cpp:308-324): float32 rows in the reference whatever the index type
  std::shared_ptr<GpuIndex> allocate_nodes(const py::array& data) {
    py::array a = as_index_dtype(data, _dt);
    if (a.ndim() != 2 || a.shape(1) != _dim) throw std::invalid_argument("Data has incorrect dimensions.");
    const int64_t n = a.shape(0);
    std::vector<int32_t> lab((size_t)n);
    std::iota(lab.begin(), lab.end(), _label_id);
    check(fnb_index_allocate_nodes(_h, a.data(), lab.data(), n));
    _label_id += (int32_t)n;
    return shared_from_this();
  }

  // PyIndex::search -> searchImpl (bindings.cpp:161-228, 337-345)
  py::tuple search(const py::array& queries, int K, int ef_search, int num_initializations) {
    py::array q = as_index_dtype(queries, _dt);
    if (q.ndim() != 2 || q.shape(1) != _dim) throw std::invalid_argument("Queries have incorrect dimensions.");
    if (K <= 0) throw std::invalid_argument("K must be positive");
    const int64_t Q = q.shape(0);
    py::array_t<float> dist({(py::ssize_t)Q, (py::ssize_t)K});
    py::array_t<int32_t> lab({(py::ssize_t)Q, (py::ssize_t)K});
    fnb_search_stats st{};
    int rc;
    {
      py::gil_scoped_release gil;  // the reference holds the GIL here; callers that thread over search() gain from letting go
      rc = fnb_search(_h, q.data(), Q, K, ef_search, num_initializations, dist.mutable_data(), lab.mutable_data(), &st);
    }
    _n_dist += (uint64_t)st.n_dist;
    check(rc);
    return py::make_tuple(dist, lab);
  }

  // PyIndex::searchSingle -> searchSingleImpl (bindings.cpp:121-159, 347-355)
  py::tuple search_single(const py::array& query, int K, int ef_search, int num_initializations) {
    py::array q = as_index_dtype(query, _dt);
    if (q.ndim() != 1 || q.shape(0) != _dim) throw std::invalid_argument("Query has incorrect dimensions.");
    if (K <= 0) throw std::invalid_argument("K must be positive");
    py::array_t<float> dist((py::ssize_t)K);
    py::array_t<int32_t> lab((py::ssize_t)K);
    fnb_search_stats st{};
    int rc;
    {
      py::gil_scoped_release gil;
      rc = fnb_search(_h, q.data(), 1, K, ef_search, num_initializations, dist.mutable_data(), lab.mutable_data(), &st);
    }
    _n_dist += (uint64_t)st.n_dist;
    check(rc);
    return py::make_tuple(dist, lab);
  }

  uint64_t get_query_distance_computations() {  // read-and-reset (bindings.cpp:270-274)
    const uint64_t n = _n_dist;
    _n_dist = 0;
    return n;
  }

  void save(const std::string& filename) { check(fnb_index_save(_h, filename.c_str())); }
  void build_graph_links(const std::string& mtx_filename) { check(fnb_index_build_graph_links(_h, mtx_filename.c_str())); }

  std::vector<std::vector<uint32_t>> get_graph_outdegree_table() {  // Index.h:240-251: self-loops are unused slots
    fnb_info info;
    check(fnb_index_info(_h, &info));
    const size_t n = info.cur_num_nodes, M = info.max_edges_per_node;
    std::vector<uint32_t> links(n * M);
    if (n) check(fnb_index_links(_h, links.data()));
    std::vector<std::vector<uint32_t>> table(n);
    for (size_t i = 0; i < n; i++)
      for (size_t j = 0; j < M; j++)
        if (links[i * M + j] != i) table[i].push_back(links[i * M + j]);
    return table;
  }

  void reorder(const std::vector<std::string>& strategies) {  // bindings.cpp:285-296 then Index.h:412-427
    for (const auto& s : strategies) {
      std::string alg = s;
      std::transform(alg.begin(), alg.end(), alg.begin(), [](unsigned char c) { return std::tolower(c); });
      if (alg != "gorder" && alg != "rcm")
        throw std::invalid_argument("`" + s + "` is not a supported graph re-ordering strategy.");
    }
    for (const auto& s : strategies) {
      if (s != "gorder" && s != "rcm") throw std::invalid_argument("Invalid reordering method: " + s);
      int rc;
      {
        py::gil_scoped_release gil;
        rc = fnb_index_reorder(_h, s == "gorder" ? FNB_REORDER_GORDER : FNB_REORDER_RCM, 5, nullptr);
      }
      check(rc);
    }
  }

  void set_num_threads(uint32_t num_threads) {  // Index.h:492-503; kept for API parity, GPU execution ignores it
    if (num_threads == 0 || num_threads > std::thread::hardware_concurrency())
      throw std::invalid_argument("Number of threads must be greater than 0 and less than or equal to the number of "
                                  "hardware threads.");
    _num_threads = num_threads;
  }
  uint32_t num_threads() const { return _num_threads; }
  uint32_t max_edges_per_node() const {
    fnb_info info;
    check(fnb_index_info(_h, &info));
    return (uint32_t)info.max_edges_per_node;
  }

  // ---- extensions of the B200 engine (no reference equivalent) ----
  py::tuple bruteforce(const py::array& queries, int K) {
    py::array q = as_index_dtype(queries, _dt);
    if (q.ndim() != 2 || q.shape(1) != _dim) throw std::invalid_argument("Queries have incorrect dimensions.");
    const int64_t Q = q.shape(0);
    py::array_t<float> dist({(py::ssize_t)Q, (py::ssize_t)K});
    py::array_t<int32_t> lab({(py::ssize_t)Q, (py::ssize_t)K});
    int rc;
    {
      py::gil_scoped_release gil;
      rc = fnb_bruteforce(_h, q.data(), Q, K, dist.mutable_data(), lab.mutable_data());
    }
    check(rc);
    return py::make_tuple(dist, lab);
  }
  py::tuple rerank(const py::array& queries, const py::array_t<int32_t, py::array::c_style | py::array::forcecast>& candidates,
                   int K, bool candidates_are_labels) {
    py::array q = as_index_dtype(queries, _dt);
    if (q.ndim() != 2 || q.shape(1) != _dim) throw std::invalid_argument("Queries have incorrect dimensions.");
    if (candidates.ndim() != 2 || candidates.shape(0) != q.shape(0) || candidates.shape(1) == 0)
      throw std::invalid_argument("candidates must be an int array of shape (num_queries, num_candidates).");
    const int64_t Q = q.shape(0);
    py::array_t<float> dist({(py::ssize_t)Q, (py::ssize_t)K});
    py::array_t<int32_t> lab({(py::ssize_t)Q, (py::ssize_t)K});
    int rc;
    {
      py::gil_scoped_release gil;
      rc = fnb_rerank(_h, q.data(), Q, candidates.data(), (int)candidates.shape(1), candidates_are_labels ? 1 : 0, K,
                      dist.mutable_data(), lab.mutable_data());
    }
    check(rc);
    return py::make_tuple(dist, lab);
  }

 protected:
  fnb_index* _h = nullptr;
  MetricType _metric;
  DataType _dt;
  bool _verbose = false;
  int _dim = 0;
  int32_t _label_id = 0;
  uint32_t _num_threads = 1;
  uint64_t _n_dist = 0;
};

// one Python class per (metric, element type), like the reference's six PyIndex specialisations
template <int METRIC, int DT>
class TypedIndex : public GpuIndex {
 public:
  using GpuIndex::GpuIndex;
  static std::shared_ptr<TypedIndex> load_index(const std::string& filename) {  // bindings.cpp:303-306
    fnb_index* h = nullptr;
    check(fnb_index_load(filename.c_str(), METRIC, DT, nullptr, 0, &h));
    return std::make_shared<TypedIndex>(h, (MetricType)METRIC, (DataType)DT, false);
  }
  static std::shared_ptr<TypedIndex> create(int dim, int dataset_size, int max_edges_per_node, bool verbose) {
    fnb_index* h = nullptr;
    check(fnb_index_create(METRIC, DT, (uint64_t)std::max(dim, 0), (uint64_t)std::max(dataset_size, 0),
                           (uint64_t)std::max(max_edges_per_node, 0), -1, &h));
    return std::make_shared<TypedIndex>(h, (MetricType)METRIC, (DataType)DT, verbose);
  }
};

template <int METRIC, int DT>
void bind_index(py::module_& m, const char* name) {
  typedef TypedIndex<METRIC, DT> T;
  py::class_<T, std::shared_ptr<T>>(m, name)
      .def("add", &T::add, py::arg("data"), py::arg("ef_construction"), py::arg("num_initializations") = 100,
           py::arg("labels") = py::none(),
           "Add vectors (num_vectors, dim) to the index; the graph is built on the GPU in batches.")
      .def("allocate_nodes", [](T& self, const py::array& data) { return std::static_pointer_cast<T>(self.allocate_nodes(data)); },
           py::arg("data"), "Append vectors as unlinked nodes (then build_graph_links).")
      .def("search_single", &T::search_single, py::arg("query"), py::arg("K"), py::arg("ef_search"),
           py::arg("num_initializations") = 100, "K nearest neighbours of one query: (distances[K], labels[K]).")
      .def("search", &T::search, py::arg("queries"), py::arg("K"), py::arg("ef_search"), py::arg("num_initializations") = 100,
           "K nearest neighbours of every query: (distances[Q, K], labels[Q, K]).")
      .def("get_query_distance_computations", &T::get_query_distance_computations,
           "Distance computations of the searches since the last call (reset on read).")
      .def("save", &T::save, py::arg("filename"), "Write the index in the reference's file format.")
      .def("build_graph_links", &T::build_graph_links, py::arg("mtx_filename"), "Fill link rows from a Matrix Market file.")
      .def("get_graph_outdegree_table", &T::get_graph_outdegree_table, "Out-links of every node.")
      .def("reorder", &T::reorder, py::arg("strategies"), "Apply `gorder` and / or `rcm` re-ordering.")
      .def("set_num_threads", &T::set_num_threads, py::arg("num_threads"), "Kept for API parity (GPU execution ignores it).")
      .def_static("load_index", &T::load_index, py::arg("filename"), "Load an index saved by flatnav or by this engine.")
      .def_property_readonly("max_edges_per_node", &T::max_edges_per_node)
      .def_property_readonly("num_threads", &T::num_threads)
      .def("bruteforce", &T::bruteforce, py::arg("queries"), py::arg("K"), "Extension: exact scan (ground truth).")
      .def("rerank", &T::rerank, py::arg("queries"), py::arg("candidates"), py::arg("K"),
           py::arg("candidates_are_labels") = true, "Extension: exact re-rank of candidate labels.");
}

template <int DT>
py::object create_typed(const std::string& distance_type, int dim, int dataset_size, int M, bool verbose) {
  if (distance_type == "l2") return py::cast(TypedIndex<FNB_METRIC_L2, DT>::create(dim, dataset_size, M, verbose));
  return py::cast(TypedIndex<FNB_METRIC_IP, DT>::create(dim, dataset_size, M, verbose));
}

}  // namespace

PYBIND11_MODULE(_core, module) {
  module.attr("__version__") = "0.2.0+b200";
  module.doc() = "flatnav on B200: the reference's Python API over libflatnav_b200.so (hand-written CUDA for sm_100a)";

  auto data_type = module.def_submodule("data_type");
  py::enum_<DataType>(data_type, "DataType")
      .value("float32", DataType::float32)
      .value("int8", DataType::int8)
      .value("uint8", DataType::uint8)
      .export_values();

  auto index = module.def_submodule("index");
  bind_index<FNB_METRIC_L2, FNB_DTYPE_FLOAT32>(index, "IndexL2Float");
  bind_index<FNB_METRIC_L2, FNB_DTYPE_INT8>(index, "IndexL2Int8");
  bind_index<FNB_METRIC_L2, FNB_DTYPE_UINT8>(index, "IndexL2Uint8");
  bind_index<FNB_METRIC_IP, FNB_DTYPE_FLOAT32>(index, "IndexIPFloat");
  bind_index<FNB_METRIC_IP, FNB_DTYPE_INT8>(index, "IndexIPInt8");
  bind_index<FNB_METRIC_IP, FNB_DTYPE_UINT8>(index, "IndexIPUint8");
  index.def(
      "create",
      [](const std::string& distance_type, int dim, int dataset_size, int max_edges_per_node, DataType index_data_type,
         bool verbose, bool /*collect_stats: distance counts are always collected here*/) -> py::object {
        std::string dt = distance_type;
        std::transform(dt.begin(), dt.end(), dt.begin(), [](unsigned char c) { return std::tolower(c); });
        if (dt != "l2" && dt != "angular")  // validateDistanceType, bindings.cpp:397-407
          throw std::invalid_argument("Invalid distance type: `" + dt + "` during index construction. Valid options "
                                      "include `l2` and `angular`.");
        switch (index_data_type) {
          case DataType::float32: return create_typed<FNB_DTYPE_FLOAT32>(distance_type, dim, dataset_size, max_edges_per_node, verbose);
          case DataType::int8: return create_typed<FNB_DTYPE_INT8>(distance_type, dim, dataset_size, max_edges_per_node, verbose);
          case DataType::uint8: return create_typed<FNB_DTYPE_UINT8>(distance_type, dim, dataset_size, max_edges_per_node, verbose);
        }
        throw std::runtime_error("Unsupported data type");
      },
      py::arg("distance_type"), py::arg("dim"), py::arg("dataset_size"), py::arg("max_edges_per_node"),
      py::arg("index_data_type") = DataType::float32, py::arg("verbose") = false, py::arg("collect_stats") = false,
      "Create an empty index (on the current CUDA device) to add() into.");

  py::enum_<MetricType>(module, "MetricType").value("L2", MetricType::L2).value("IP", MetricType::IP);
}
