"""`flatnav.index` of the reference, search side, on the B200 engine.

Mirrors the classes `IndexL2Float / IndexL2Uint8 / IndexL2Int8 / IndexIPFloat / IndexIPUint8 / IndexIPInt8`
bound in python-bindings/src/flatnav/bindings.cpp:358-395, 426-474 of the reference: same method names,
argument meaning, return dtypes/shapes and exception types for `load_index`, `search`, `search_single`,
`save`, `set_num_threads`, `num_threads`, `max_edges_per_node`, `get_query_distance_computations`, and — the caller
side of the search path, SURVEY.md §8f — `create` / `add` (GPU batched construction, csrc/build.cu).
`allocate_nodes` / `build_graph_links` (Matrix Market link import), `get_graph_outdegree_table` and `reorder`
(gorder / rcm, csrc/reorder.cu) complete the binding's method list.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi
from .data_type import DataType

_NP = {DataType.float32: np.float32, DataType.uint8: np.uint8, DataType.int8: np.int8}

_from_buffer = C.c_char.from_buffer
_addressof = C.addressof


def _ptr(a: np.ndarray) -> int:
    """Address of a C-contiguous array's data.  `ndarray.ctypes.data` builds a helper object on every access (~1.9 us,
    three of them per search call: as much as everything else the wrapper does for one query); the buffer protocol gives
    the address in ~0.3 us, for writable non-empty arrays."""
    if a.size and a.flags.writeable:
        return _addressof(_from_buffer(a))
    return a.ctypes.data


class _GpuIndex:
    _metric: int = _capi.FNB_METRIC_L2
    _data_type: DataType = DataType.float32

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)
        info = _capi.FnbInfo()
        _capi.check(_capi.lib().fnb_index_info(self._h, C.byref(info)))
        self._info = info
        self._dim = int(info.dim)
        # Index::loadIndex sets _num_threads = max(1, hardware_concurrency / 2)  (Index.h:467)
        self._num_threads = max(1, (os.cpu_count() or 1) // 2)
        self._n_dist = 0
        self._label_id = 0  # PyIndex::_label_id (bindings.cpp:232): labels handed out by allocate_nodes
        self._last_st = None  # FnbSearchStats of the last search call

    @property
    def last_stats(self) -> dict:
        """Counters of the last `search` / `search_single` call (n_queries, n_dist, n_hops, kernel_ms, ...)."""
        return self._last_st.as_dict() if self._last_st is not None else {}

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _capi.lib().fnb_index_free(h)
            except Exception:
                pass

    # ---- loading / saving -------------------------------------------------------------------
    @classmethod
    def load_index(cls, filename: str, devices=None):
        """Index<dist_t,label_t>::loadIndex (Index.h:442-479; bindings.cpp:303-306, :471).

        `devices` (extension): CUDA device ids to replicate the index on; queries of a batch are then
        split evenly across the replicas.  Default: the current device."""
        out = C.c_void_p()
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            rc = _capi.lib().fnb_index_load(os.fsencode(filename), cls._metric, int(cls._data_type), arr, len(devices),
                                            C.byref(out))
        else:
            rc = _capi.lib().fnb_index_load(os.fsencode(filename), cls._metric, int(cls._data_type), None, 0,
                                            C.byref(out))
        _capi.check(rc)
        return cls(out.value)

    @classmethod
    def from_bytes(cls, blob, devices=None):
        buf = np.frombuffer(blob, dtype=np.uint8)
        out = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices) if devices else None
        _capi.check(_capi.lib().fnb_index_from_memory(buf.ctypes.data, buf.size, cls._metric, int(cls._data_type), arr,
                                                      len(devices) if devices else 0, C.byref(out)))
        return cls(out.value)

    def save(self, filename: str) -> None:
        """Index::saveIndex (Index.h:481-490): writes the reference's cereal byte layout."""
        _capi.check(_capi.lib().fnb_index_save(self._h, os.fsencode(filename)))

    # ---- search -----------------------------------------------------------------------------
    def _cast(self, a) -> np.ndarray:
        # py::array_t<T, c_style | forcecast>  (bindings.cpp:36-51)
        return np.ascontiguousarray(a, dtype=_NP[self._data_type])

    def search(self, queries, K: int, ef_search: int, num_initializations: int = 100, *, out=None,
               exact_rerank: bool = False):
        """PyIndex::search -> searchImpl (bindings.cpp:337-345, 161-228).

        Returns (distances float32 [Q,K], labels int32 [Q,K]).  `out` (extension): a pair of preallocated
        C-contiguous arrays of those shapes/dtypes (e.g. pinned memory) to write into instead of fresh ones.
        `exact_rerank` (extension): take the whole max(ef_search, K)-long candidate list of the
        traversal and re-rank it with `rerank` before cutting to K.  The traversal already evaluates exact distances,
        so this returns what the plain call returns (a property the tests pin); it exists for pipelines whose first
        stage is approximate and for protocol parity with drivers that re-rank."""
        q = self._cast(queries)
        if q.ndim != 2 or q.shape[1] != self._dim:
            raise ValueError("Queries have incorrect dimensions.")
        if exact_rerank:
            _, cand = self.search(q, max(int(ef_search), int(K)), ef_search, num_initializations)
            d, l = self.rerank(q, cand, K)
            if out is not None:
                out[0][...], out[1][...] = d, l
                return out
            return d, l
        Q = q.shape[0]
        if out is not None:
            dist, lab = out
            if (dist.shape != (Q, K) or lab.shape != (Q, K) or dist.dtype != np.float32 or lab.dtype != np.int32
                    or not dist.flags.c_contiguous or not lab.flags.c_contiguous):
                raise ValueError("out must be C-contiguous (float32[Q,K], int32[Q,K])")
        else:
            dist = np.empty((Q, K), dtype=np.float32)
            lab = np.empty((Q, K), dtype=np.int32)
        st = _capi.FnbSearchStats()
        rc = _capi.lib().fnb_search(self._h, _ptr(q), Q, int(K), int(ef_search), int(num_initializations),
                                    _ptr(dist), _ptr(lab), C.byref(st))
        self._last_st = st
        self._n_dist += int(st.n_dist)
        _capi.check(rc)
        return dist, lab

    def search_single(self, query, K: int, ef_search: int, num_initializations: int = 100):
        """PyIndex::searchSingle -> searchSingleImpl (bindings.cpp:347-355, 121-159).

        Returns (distances float32 [K], labels int32 [K])."""
        q = self._cast(query)
        if q.ndim != 1 or q.shape[0] != self._dim:
            raise ValueError("Query has incorrect dimensions.")
        d, l = self.search(q[None, :], K, ef_search, num_initializations)
        return d[0], l[0]

    def bruteforce(self, queries, K: int):
        """Exact scan (extension; ground truth / exact re-rank). Returns (distances, labels) like search()."""
        q = self._cast(queries)
        if q.ndim != 2 or q.shape[1] != self._dim:
            raise ValueError("Queries have incorrect dimensions.")
        Q = q.shape[0]
        dist = np.empty((Q, K), dtype=np.float32)
        lab = np.empty((Q, K), dtype=np.int32)
        rc = _capi.lib().fnb_bruteforce(self._h, q.ctypes.data, Q, int(K), dist.ctypes.data, lab.ctypes.data)
        st = _capi.FnbBfStats()
        _capi.lib().fnb_bruteforce_stats(C.byref(st))
        self.last_bruteforce_stats = st.as_dict()
        _capi.check(rc)
        return dist, lab

    def rerank(self, queries, candidates, K: int, candidates_are_labels: bool = True):
        """Exact re-rank (extension): for every query the candidate labels (or node ids) of
        `candidates[q]` (int array [Q, C]) are evaluated exactly and the K best by (distance, node id) returned as
        (distances float32 [Q,K], labels int32 [Q,K]); unknown / negative / repeated candidates are skipped, unfilled
        slots are +inf / -1."""
        q = self._cast(queries)
        if q.ndim != 2 or q.shape[1] != self._dim:
            raise ValueError("Queries have incorrect dimensions.")
        c = np.ascontiguousarray(candidates, dtype=np.int32)
        if c.ndim != 2 or c.shape[0] != q.shape[0] or c.shape[1] == 0:
            raise ValueError("candidates must be an int array of shape (num_queries, num_candidates).")
        Q = q.shape[0]
        dist = np.empty((Q, K), dtype=np.float32)
        lab = np.empty((Q, K), dtype=np.int32)
        _capi.check(_capi.lib().fnb_rerank(self._h, q.ctypes.data, Q, c.ctypes.data, c.shape[1],
                                           1 if candidates_are_labels else 0, int(K), dist.ctypes.data, lab.ctypes.data))
        return dist, lab

    def search_device(self, d_queries: int, Q: int, K: int, ef_search: int, num_initializations: int, d_out_dist: int,
                      d_out_label: int, stream: int = 0, d_ndist: int = 0, d_nhops: int = 0, replica: int = 0) -> None:
        """Kernel-only path on device-resident buffers (raw device pointers as ints); asynchronous."""
        _capi.check(_capi.lib().fnb_search_device(self._h, replica, d_queries, Q, K, ef_search, num_initializations,
                                                  d_out_dist, d_out_label, d_ndist or None, d_nhops or None,
                                                  stream or None))

    def device_totals(self, replica: int = 0):
        nd, nh, ns = C.c_int64(), C.c_int64(), C.c_int64()
        _capi.check(_capi.lib().fnb_search_device_totals(self._h, replica, C.byref(nd), C.byref(nh), C.byref(ns)))
        return nd.value, nh.value, ns.value

    def kernel_signature(self, Q: int, K: int, ef_search: int) -> str:
        """Name of the traversal-kernel instantiation a search of this shape launches (as ncu prints it, no blanks)."""
        buf = C.create_string_buffer(128)
        _capi.check(_capi.lib().fnb_search_kernel_signature(self._h, int(Q), int(K), int(ef_search), buf, 128))
        return buf.value.decode()

    def search_plan(self, Q: int, K: int, ef_search: int) -> dict:
        """Launch plan of a search of this shape: shared memory per query, visited-set slots, resident CTAs per SM."""
        pi = _capi.FnbPlanInfo()
        _capi.check(_capi.lib().fnb_search_plan(self._h, int(Q), int(K), int(ef_search), C.byref(pi)))
        return pi.as_dict()

    # ---- getters / knobs --------------------------------------------------------------------
    def get_query_distance_computations(self) -> int:
        """PyIndex::getQueryDistanceComputations (bindings.cpp:270-274): read-and-reset.

        Counts database-row distance evaluations made by search() calls since the last read, including the
        entry-selection probes (the reference adds `num_initializations` per search instead, Index.h:857-859,
        and counts nothing at all on a loaded index because `collect_stats` is not serialised)."""
        n, self._n_dist = self._n_dist, 0
        return n

    def set_num_threads(self, num_threads: int) -> None:
        """Index::setNumThreads (Index.h:492-503). Kept for API parity; GPU execution ignores it."""
        if num_threads <= 0 or num_threads > (os.cpu_count() or 1):
            raise ValueError("Number of threads must be greater than 0 and less than or equal to "
                             "the number of hardware threads.")
        self._num_threads = int(num_threads)

    @property
    def num_threads(self) -> int:
        return self._num_threads

    @property
    def max_edges_per_node(self) -> int:
        return int(self._info.max_edges_per_node)

    @property
    def info(self) -> dict:
        d = {k: getattr(self._info, k) for k, _ in self._info._fields_ if k != "device_ids"}
        d["device_ids"] = list(self._info.device_ids)[: self._info.n_devices]
        return d

    # ---- construction --------------------------------------------------------------------------
    def add(self, data, ef_construction: int, num_initializations: int = 100, labels=None) -> None:
        """PyIndex::add -> addImpl (bindings.cpp:62-110, :436-443) -> Index::addBatch (Index.h:301-330).

        `data`: array castable to the index dtype, shape (num_vectors, dim); `labels`: optional sequence of
        num_vectors ints (default 0 .. num_vectors-1, as in the binding)."""
        a = np.asarray(data)
        if a.ndim != 2 or a.shape[1] != self._dim:
            nd, dd = a.ndim, (a.shape[1] if a.ndim > 1 else -1)
            raise ValueError(f"Data has incorrect dimensions. data.ndim() = `{nd}` and data_dim = `{dd}`. Expected 2D "
                             "array with dimensions (num_vectors, dim).")
        a = self._cast(a)
        lab = None
        if labels is not None:
            try:
                lab = np.ascontiguousarray(np.asarray(list(labels)), dtype=np.int32)
            except (TypeError, ValueError):
                raise ValueError("Invalid labels provided.")
            if lab.ndim != 1 or lab.shape[0] != a.shape[0]:
                raise ValueError("Incorrect number of labels.")
        st = _capi.FnbBuildStats()
        rc = _capi.lib().fnb_index_add(self._h, a.ctypes.data, lab.ctypes.data if lab is not None else None,
                                       a.shape[0], int(ef_construction), int(num_initializations), C.byref(st))
        self.last_build_stats = st.as_dict()
        _capi.check(rc)
        _capi.check(_capi.lib().fnb_index_info(self._h, C.byref(self._info)))

    def reserve(self, max_node_count: int) -> None:
        """Extension: make room for more nodes (a loaded index holds exactly its current node count)."""
        _capi.check(_capi.lib().fnb_index_reserve(self._h, int(max_node_count)))
        _capi.check(_capi.lib().fnb_index_info(self._h, C.byref(self._info)))

    def allocate_nodes(self, data):
        """PyIndex::allocateNodes (bindings.cpp:308-324, :441-446): appends the rows of `data` as unlinked nodes
        (every link slot a self-loop, Index.h:262-272), labelled by a running counter that starts at 0, and returns
        the index so that `.build_graph_links(...)` can be chained.  The binding takes float32 rows and copies their
        bytes whatever the index type; here they are cast to the index data type."""
        a = np.asarray(data)
        if a.ndim != 2 or a.shape[1] != self._dim:
            raise ValueError("Data has incorrect dimensions.")
        a = self._cast(a)
        lab = np.arange(self._label_id, self._label_id + a.shape[0], dtype=np.int32)
        _capi.check(_capi.lib().fnb_index_allocate_nodes(self._h, a.ctypes.data, lab.ctypes.data, a.shape[0]))
        self._label_id += a.shape[0]
        _capi.check(_capi.lib().fnb_index_info(self._h, C.byref(self._info)))
        return self

    def build_graph_links(self, mtx_filename: str) -> None:
        """PyIndex::buildGraphLinks -> Index::buildGraphLinks (bindings.cpp:276-278, Index.h:187-238): fills the link
        rows of allocated nodes from a Matrix Market edge list."""
        _capi.check(_capi.lib().fnb_index_build_graph_links(self._h, os.fsencode(mtx_filename)))

    def get_graph_outdegree_table(self):
        """PyIndex::getGraphOutdegreeTable -> Index::getGraphOutdegreeTable (bindings.cpp:281, Index.h:240-251):
        for every node the list of its out-links, self-loops (unused slots) removed."""
        n, M = int(self._info.cur_num_nodes), int(self._info.max_edges_per_node)
        links = np.empty((n, M), dtype=np.uint32)
        _capi.check(_capi.lib().fnb_index_links(self._h, links.ctypes.data))
        own = np.arange(n, dtype=np.uint32)[:, None]
        return [row[keep].tolist() for row, keep in zip(links, links != own)]

    def reorder(self, strategies) -> None:
        """PyIndex::reorder -> Index::doGraphReordering (bindings.cpp:285-296, Index.h:412-427): applies `gorder`
        (window 5) and / or `rcm` in the given sequence.  Same validation as the binding: every name is checked
        case-insensitively first; the index then matches them case-sensitively, so `"GORDER"` passes the first check
        and fails the second, after the strategies before it have been applied — as in the reference.
        The last permutation (old node id -> new node id) is kept in `last_permutation`."""
        strategies = list(strategies)
        for st in strategies:
            if str(st).lower() not in ("gorder", "rcm"):
                raise ValueError("`" + str(st) + "` is not a supported graph re-ordering strategy.")
        for st in strategies:
            if st not in ("gorder", "rcm"):
                raise ValueError("Invalid reordering method: " + str(st))
            perm = np.empty(int(self._info.cur_num_nodes), dtype=np.uint32)
            method = _capi.FNB_REORDER_GORDER if st == "gorder" else _capi.FNB_REORDER_RCM
            _capi.check(_capi.lib().fnb_index_reorder(self._h, method, 5, perm.ctypes.data))
            self.last_permutation = perm

    def relabel(self, permutation) -> None:
        """Index::relabel (Index.h:872-926) with a caller-supplied permutation: node i moves to row permutation[i]."""
        perm = np.ascontiguousarray(permutation, dtype=np.uint32)
        if perm.ndim != 1 or perm.shape[0] != int(self._info.cur_num_nodes):
            raise ValueError("permutation must have one entry per node")
        _capi.check(_capi.lib().fnb_index_relabel(self._h, perm.ctypes.data))


class IndexL2Float(_GpuIndex):
    _metric, _data_type = _capi.FNB_METRIC_L2, DataType.float32


class IndexL2Uint8(_GpuIndex):
    _metric, _data_type = _capi.FNB_METRIC_L2, DataType.uint8


class IndexL2Int8(_GpuIndex):
    _metric, _data_type = _capi.FNB_METRIC_L2, DataType.int8


class IndexIPFloat(_GpuIndex):
    _metric, _data_type = _capi.FNB_METRIC_IP, DataType.float32


class IndexIPUint8(_GpuIndex):
    _metric, _data_type = _capi.FNB_METRIC_IP, DataType.uint8


class IndexIPInt8(_GpuIndex):
    _metric, _data_type = _capi.FNB_METRIC_IP, DataType.int8


_CLASSES = {
    ("l2", DataType.float32): IndexL2Float, ("l2", DataType.uint8): IndexL2Uint8, ("l2", DataType.int8): IndexL2Int8,
    ("angular", DataType.float32): IndexIPFloat, ("angular", DataType.uint8): IndexIPUint8,
    ("angular", DataType.int8): IndexIPInt8,
}


def index_class(distance_type: str, index_data_type: DataType = DataType.float32):
    """Class the reference's `create(distance_type, ..., index_data_type)` would instantiate (bindings.cpp:409-424)."""
    dt = distance_type.lower()
    if dt not in ("l2", "angular"):
        raise ValueError("Invalid distance type: `" + dt + "` during index construction. Valid options "
                         "include `l2` and `angular`.")  # validateDistanceType, bindings.cpp:397-407
    return _CLASSES[(dt, DataType(index_data_type))]


def create(distance_type: str, dim: int, dataset_size: int, max_edges_per_node: int,
           index_data_type: DataType = DataType.float32, verbose: bool = False, collect_stats: bool = False):
    """`flatnav.index.create` (bindings.cpp:484-504): an empty index on the current CUDA device to `add()` into.
    `verbose` prints the parameters like Index::getIndexSummary; `collect_stats` is accepted (distance counts are
    always collected here)."""
    cls = index_class(distance_type, index_data_type)
    out = C.c_void_p()
    _capi.check(_capi.lib().fnb_index_create(cls._metric, int(cls._data_type), int(dim), int(dataset_size),
                                             int(max_edges_per_node), -1, C.byref(out)))
    ix = cls(out.value)
    if verbose:
        print(f"max_edges_per_node: {max_edges_per_node}\nmax_node_count: {dataset_size}\ndimension: {dim}", flush=True)
    return ix
