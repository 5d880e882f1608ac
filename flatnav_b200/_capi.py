"""ctypes binding of libflatnav_b200.so (include/flatnav_b200.h).

The shared library is built in-tree by `flatnav_b200/csrc/Makefile` (see `__graft_entry__.build`).
There is no fallback: if the library is missing the import of the search path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FNB_LIB_PATH") or os.path.join(HERE, "libflatnav_b200.so")  # env: development A/B builds

FNB_OK, FNB_SHORT_RESULT = 0, 1
FNB_ERR_INVALID_ARG, FNB_ERR_IO, FNB_ERR_FORMAT, FNB_ERR_CUDA, FNB_ERR_UNSUPPORTED, FNB_ERR_NOMEM = -1, -2, -3, -4, -5, -6
FNB_DTYPE_UINT8, FNB_DTYPE_INT8, FNB_DTYPE_FLOAT32, FNB_DTYPE_ANY = 0, 4, 9, -1
FNB_METRIC_L2, FNB_METRIC_IP = 0, 1
FNB_IPC_HANDLE_BYTES = 64
FNB_REORDER_GORDER, FNB_REORDER_RCM = 0, 1


class FnbInfo(C.Structure):
    _fields_ = [
        ("data_type", C.c_int32), ("metric", C.c_int32),
        ("max_edges_per_node", C.c_uint64), ("dim", C.c_uint64), ("data_size_bytes", C.c_uint64),
        ("node_size_bytes", C.c_uint64), ("max_node_count", C.c_uint64), ("cur_num_nodes", C.c_uint64),
        ("n_devices", C.c_int32), ("device_ids", C.c_int32 * 16), ("device_bytes", C.c_uint64),
        ("row_stride_bytes", C.c_uint32), ("lanes_per_row", C.c_uint32),
    ]


class FnbSearchStats(C.Structure):
    _fields_ = [
        ("n_queries", C.c_int64), ("n_dist", C.c_int64), ("n_hops", C.c_int64), ("n_short", C.c_int64),
        ("algo_bytes", C.c_int64), ("kernel_ms", C.c_float), ("total_ms", C.c_float),
        ("kernel_launches", C.c_int32), ("reserved", C.c_int32),
    ]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class FnbPlanInfo(C.Structure):
    _fields_ = [("latency_variant", C.c_int32), ("dense_plan", C.c_int32), ("list_capacity", C.c_int32),
                ("visited_slots", C.c_int32), ("smem_bytes_per_query", C.c_int32), ("ctas_per_sm", C.c_int32),
                ("warps_per_sm", C.c_int32), ("queries_per_sm", C.c_int32)]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class FnbBuildStats(C.Structure):
    _fields_ = [("n_added", C.c_int64), ("n_batches", C.c_int64), ("n_dropped_backlinks", C.c_int64),
                ("device_ms", C.c_float), ("reserved", C.c_float)]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class FnbBfStats(C.Structure):
    _fields_ = [
        ("path", C.c_int32), ("reserved", C.c_int32), ("n_unsafe", C.c_int64), ("n_candidates", C.c_int64),
        ("prep_ms", C.c_float), ("gemm_ms", C.c_float), ("rerank_ms", C.c_float), ("rescan_ms", C.c_float),
        ("gemm_flops", C.c_double),
    ]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


EXPORTS = {
    # name: (restype, argtypes)
    "fnb_index_load": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "fnb_index_from_memory": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int,
                                        C.POINTER(C.c_void_p)]),
    "fnb_index_create": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]),
    "fnb_index_reserve": (C.c_int, [C.c_void_p, C.c_uint64]),
    "fnb_index_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                C.POINTER(FnbBuildStats)]),
    "fnb_index_allocate_nodes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "fnb_index_build_graph_links": (C.c_int, [C.c_void_p, C.c_char_p]),
    "fnb_index_links": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fnb_index_reorder": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "fnb_graph_order": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_void_p]),
    "fnb_index_relabel": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fnb_index_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "fnb_index_info": (C.c_int, [C.c_void_p, C.POINTER(FnbInfo)]),
    "fnb_index_free": (None, [C.c_void_p]),
    "fnb_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                             C.POINTER(FnbSearchStats)]),
    "fnb_search_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fnb_search_device_totals": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                           C.POINTER(C.c_int64)]),
    "fnb_search_kernel_signature": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_char_p, C.c_size_t]),
    "fnb_search_plan": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(FnbPlanInfo)]),
    "fnb_bruteforce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "fnb_rerank": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fnb_bruteforce_stats": (C.c_int, [C.POINTER(FnbBfStats)]),
    "fnb_merge_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "fnb_exchange_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    "fnb_exchange_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fnb_exchange_attach": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fnb_search_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "fnb_exchange_status": (C.c_int, [C.c_void_p]),
    "fnb_exchange_free": (None, [C.c_void_p]),
    "fnb_last_error": (C.c_char_p, []),
    "fnb_version": (C.c_char_p, []),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C flatnav_b200/csrc`). flatnav_b200 has no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            try:
                fn = getattr(l, name)  # AttributeError if the library does not export a declared symbol
            except AttributeError:
                if os.environ.get("FNB_LIB_PATH"):  # development A/B against an older build: tolerate newer symbols
                    continue
                raise
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def last_error() -> str:
    msg = lib().fnb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int) -> int:
    """Map a status code to the exception the reference's binding raises for the same condition."""
    if rc == FNB_OK:
        return rc
    msg = last_error()
    if rc == FNB_SHORT_RESULT:
        raise RuntimeError(msg)  # bindings.cpp:134-137, 184-189
    if rc == FNB_ERR_INVALID_ARG:
        raise ValueError(msg)  # std::invalid_argument -> ValueError through pybind11
    if rc == FNB_ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)  # std::runtime_error
