"""TEST INFRASTRUCTURE — ctypes wrapper over oracle/liboracle.so (the CPU restatement).

See flatnav_oracle.cpp for the reference file:line each function follows.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")

F32, U8, I8 = 9, 0, 4  # flatnav::util::DataType values (util/Datatype.h:11-24)
L2, IP = 0, 1
NP_DTYPE = {F32: np.float32, U8: np.uint8, I8: np.int8}
MODE_HEAPS, MODE_LIST = 0, 1
ORDER_SEQUENTIAL, ORDER_LANES = 0, 1


class _OraIndex(C.Structure):
    _fields_ = [
        ("data_type", C.c_int32),
        ("metric", C.c_int32),
        ("M", C.c_uint64),
        ("data_size_bytes", C.c_uint64),
        ("node_size_bytes", C.c_uint64),
        ("max_node_count", C.c_uint64),
        ("cur_num_nodes", C.c_uint64),
        ("dim", C.c_uint64),
        ("mem", C.c_void_p),
    ]


_lib = None


def build() -> None:
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.ora_parse_header.restype = C.c_int
        _lib.ora_parse_header.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(_OraIndex)]
        _lib.ora_search.restype = C.c_int64
        _lib.ora_search.argtypes = [C.POINTER(_OraIndex), C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ora_distance.restype = C.c_float
        _lib.ora_distance.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64]
        _lib.ora_bruteforce.restype = C.c_int64
        _lib.ora_bruteforce.argtypes = [C.POINTER(_OraIndex), C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p]
        _lib.ora_reorder.restype = C.c_int
        _lib.ora_reorder.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
        _lib.ora_build_graph_links.restype = C.c_int
        _lib.ora_build_graph_links.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64]
    return _lib


GORDER, RCM = 0, 1


def reorder_file(blob, strategies, window: int = 5):
    """Index::doGraphReordering (Index.h:412-427) on the bytes of an index file.  Returns (new file bytes as a
    uint8 array, list of permutations old id -> new id, one per strategy)."""
    buf = np.array(np.frombuffer(blob, dtype=np.uint8))  # writable copy
    n = int(np.frombuffer(buf[4 + 32:4 + 40].tobytes(), dtype=np.uint64)[0])  # cur_num_nodes
    perms = []
    for st in strategies:
        perm = np.empty(n, dtype=np.uint32)
        rc = lib().ora_reorder(buf.ctypes.data, buf.size, {"gorder": GORDER, "rcm": RCM}[st], window, perm.ctypes.data)
        if rc != 0:
            raise ValueError(f"ora_reorder failed with {rc}")
        perms.append(perm)
    return buf, perms


def build_graph_links_file(blob, src, dst):
    """Edge rule of Index::buildGraphLinks (Index.h:219-234) on the bytes of an index file; 0-based edges."""
    buf = np.array(np.frombuffer(blob, dtype=np.uint8))
    src = np.ascontiguousarray(src, dtype=np.uint32)
    dst = np.ascontiguousarray(dst, dtype=np.uint32)
    rc = lib().ora_build_graph_links(buf.ctypes.data, buf.size, src.ctypes.data, dst.ctypes.data, src.size)
    if rc != 0:
        raise ValueError(f"ora_build_graph_links failed with {rc}")
    return buf


class OracleIndex:
    """An index file in the reference's cereal layout, searched by the CPU restatement."""

    def __init__(self, path_or_bytes, metric: int):
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview, np.ndarray)):
            self._buf = np.frombuffer(path_or_bytes, dtype=np.uint8)
        else:
            self._buf = np.fromfile(path_or_bytes, dtype=np.uint8)
        self._ix = _OraIndex()
        rc = lib().ora_parse_header(self._buf.ctypes.data, self._buf.size, metric, C.byref(self._ix))
        if rc != 0:
            raise ValueError(f"ora_parse_header failed with {rc}")
        self.metric = metric
        self.data_type = self._ix.data_type
        self.M = int(self._ix.M)
        self.dim = int(self._ix.dim)
        self.data_size_bytes = int(self._ix.data_size_bytes)
        self.node_size_bytes = int(self._ix.node_size_bytes)
        self.max_node_count = int(self._ix.max_node_count)
        self.cur_num_nodes = int(self._ix.cur_num_nodes)
        self.np_dtype = NP_DTYPE[self.data_type]

    # views into the AoS blob -------------------------------------------------------------
    def _nodes(self) -> np.ndarray:
        return self._buf[60:60 + self.node_size_bytes * self.max_node_count].reshape(self.max_node_count,
                                                                                  self.node_size_bytes)

    def vectors(self) -> np.ndarray:
        raw = np.ascontiguousarray(self._nodes()[: self.cur_num_nodes, : self.data_size_bytes])
        return raw.view(self.np_dtype).reshape(self.cur_num_nodes, self.dim)

    def links(self) -> np.ndarray:
        raw = np.ascontiguousarray(
            self._nodes()[: self.cur_num_nodes, self.data_size_bytes: self.data_size_bytes + 4 * self.M])
        return raw.view(np.uint32).reshape(self.cur_num_nodes, self.M)

    def labels(self) -> np.ndarray:
        raw = np.ascontiguousarray(self._nodes()[: self.cur_num_nodes, self.data_size_bytes + 4 * self.M:])
        return raw.view(np.int32).reshape(self.cur_num_nodes)

    # search ------------------------------------------------------------------------------
    def search(self, queries: np.ndarray, K: int, ef_search: int, num_initializations: int = 100, *,
               mode: int = MODE_HEAPS, dist_order: int = ORDER_LANES, threads: int = 1, counters: bool = False):
        q = np.ascontiguousarray(queries, dtype=self.np_dtype)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError("Queries have incorrect dimensions.")
        Q = q.shape[0]
        d = np.empty((Q, K), dtype=np.float32)
        l = np.empty((Q, K), dtype=np.int32)
        nd = np.zeros(Q, dtype=np.int64)
        nh = np.zeros(Q, dtype=np.int64)
        rc = lib().ora_search(C.byref(self._ix), q.ctypes.data, Q, K, ef_search, num_initializations, mode,
                              dist_order, threads, d.ctypes.data, l.ctypes.data, nd.ctypes.data, nh.ctypes.data)
        if rc == -10:
            raise ValueError("num_initializations must be greater than 0.")
        if rc < 0:
            raise RuntimeError(f"ora_search failed with {rc}")
        if counters:
            return d, l, nd, nh
        return d, l

    def bruteforce(self, queries: np.ndarray, K: int, *, dist_order: int = ORDER_LANES, threads: int = 0):
        q = np.ascontiguousarray(queries, dtype=self.np_dtype)
        Q = q.shape[0]
        d = np.empty((Q, K), dtype=np.float32)
        l = np.empty((Q, K), dtype=np.int32)
        threads = threads or (os.cpu_count() or 1)
        lib().ora_bruteforce(C.byref(self._ix), q.ctypes.data, Q, K, dist_order, threads, d.ctypes.data,
                             l.ctypes.data)
        return d, l


def distance(x: np.ndarray, y: np.ndarray, metric: int, dist_order: int = ORDER_LANES) -> float:
    dt = {np.dtype(np.float32): F32, np.dtype(np.uint8): U8, np.dtype(np.int8): I8}[x.dtype]
    x = np.ascontiguousarray(x)
    y = np.ascontiguousarray(y)
    return float(lib().ora_distance(dt, metric, dist_order, x.ctypes.data, y.ctypes.data, x.size))
