"""CPU: the C-ABI library loads, exports every symbol include/flatnav_b200.h declares, and the host-side
mirror of the reference's Python interface has the reference's names / validation / errors.  No compute."""
import os
import re

import numpy as np
import pytest

import flatnav_b200
from conftest import ROOT, golden_index_path
from flatnav_b200 import _capi
from flatnav_b200.data_type import DataType


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "flatnav_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(fnb_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = _capi.lib()
    syms = declared_symbols()
    assert len(syms) >= 11
    for s in syms:
        assert hasattr(lib, s), s
    assert set(syms) <= set(_capi.EXPORTS) | {"fnb_search_device_totals"}
    assert b"sm_100a" in lib.fnb_version()


def test_namespace_mirrors_reference():
    # python-bindings/src/flatnav/__init__.py:9-27 and bindings.cpp:358-395, 507-521
    for name in ("IndexL2Float", "IndexIPFloat", "IndexL2Uint8", "IndexIPUint8", "IndexL2Int8", "IndexIPInt8", "create"):
        assert hasattr(flatnav_b200.index, name)
    assert int(DataType.float32) == 9 and int(DataType.int8) == 4 and int(DataType.uint8) == 0
    assert flatnav_b200.MetricType.L2 == 0 and flatnav_b200.MetricType.IP == 1
    for m in ("search", "search_single", "load_index", "save", "set_num_threads", "get_query_distance_computations",
              "add", "allocate_nodes", "reorder", "build_graph_links", "get_graph_outdegree_table"):
        assert hasattr(flatnav_b200.index.IndexL2Float, m)


def test_create_validates_distance_type_like_reference():
    with pytest.raises(ValueError, match="Invalid distance type"):  # bindings.cpp:397-407
        flatnav_b200.index.create("cosine", 8, 10, 4)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        with pytest.raises(RuntimeError):  # no CUDA device and no CPU path: loud failure, not a fallback
            flatnav_b200.index.create("l2", 8, 10, 4)
    for name in ("reorder", "allocate_nodes", "build_graph_links", "get_graph_outdegree_table"):
        assert callable(getattr(flatnav_b200.index.IndexL2Float, name))  # the binding's method list, bindings.cpp:436-473
    assert flatnav_b200.index.index_class("angular", DataType.uint8) is flatnav_b200.index.IndexIPUint8
    assert flatnav_b200.index.index_class("L2") is flatnav_b200.index.IndexL2Float


def test_load_missing_file_is_runtime_error():
    with pytest.raises(RuntimeError, match="Unable to open file for reading"):  # Index.h:445-447
        flatnav_b200.index.IndexL2Float.load_index("/nonexistent/file.idx")


def test_header_validation_happens_before_any_device_work(tmp_path):
    raw = bytearray(open(golden_index_path("l2_f32_d7"), "rb").read())
    # wrong dtype requested for the class
    with pytest.raises(RuntimeError, match="data_type"):
        flatnav_b200.index.IndexL2Uint8.from_bytes(bytes(raw))
    bad = bytearray(raw)
    bad[20:28] = (12345).to_bytes(8, "little")  # node_size != data_size + 4M + 4
    with pytest.raises(RuntimeError, match="node_size"):
        flatnav_b200.index.IndexL2Float.from_bytes(bytes(bad))
    with pytest.raises(RuntimeError, match="truncated"):
        flatnav_b200.index.IndexL2Float.from_bytes(bytes(raw[:1000]))
    with pytest.raises(RuntimeError, match="60-byte"):
        flatnav_b200.index.IndexL2Float.from_bytes(bytes(raw[:10]))


def test_no_cpu_fallback_without_device():
    import ctypes as C
    n = C.c_int(0)
    try:
        cudart = C.CDLL("libcudart.so.12")
        have = cudart.cudaGetDeviceCount(C.byref(n)) == 0 and n.value > 0
    except OSError:
        have = False
    if have:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        flatnav_b200.index.IndexL2Float.load_index(golden_index_path("l2_f32_d7"))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "flatnav_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_cpp_shim_compiles_and_links(tmp_path):
    """the C++ surface (include/flatnav_b200/Index.h) builds against the C ABI with plain g++"""
    import subprocess
    exe = str(tmp_path / "shim_test")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "shim_test.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "flatnav_b200"), "-lflatnav_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "flatnav_b200")], check=True)
    r = subprocess.run([exe], capture_output=True)
    assert r.returncode == 2  # usage


def _header(dt=9, M=32, data_size=512, node_size=644, max_nodes=10, cur=10, dim=128, ds2=512):
    return np.int32(dt).tobytes() + np.array([M, data_size, node_size, max_nodes, cur, dim, ds2], dtype=np.uint64).tobytes()


@pytest.mark.parametrize("kw", [
    dict(M=2**62 + 8, node_size=(512 + 4 * (2**62 + 8) + 4) % 2**64),      # 4*M wraps: node_size check would pass
    dict(M=70000, node_size=512 + 4 * 70000 + 4),                          # beyond what the kernels index
    dict(dim=2**62, data_size=0, ds2=0, node_size=4 * 32 + 4),             # dim * 4 wraps to 0
    dict(max_nodes=2**40, cur=2**40),                                      # node count beyond uint32 node ids
    dict(node_size=2**63, max_nodes=2**31 - 1),                            # node_size * max_nodes would wrap
    dict(cur=11),                                                          # cur_num_nodes > max_node_count
    dict(dt=3),                                                            # unknown data type
], ids=["M-wraps", "M-huge", "dim-wraps", "too-many-nodes", "blob-wraps", "cur>max", "dtype"])
def test_hostile_headers_are_rejected_before_any_allocation(kw):
    """a crafted header must fail the checks of parse_header, not wrap its size arithmetic (runs without a device:
    the header is parsed before the device is touched)"""
    import ctypes as C
    blob = _header(**kw) + bytes(644 * 10)
    out = C.c_void_p()
    rc = _capi.lib().fnb_index_from_memory(blob, len(blob), _capi.FNB_METRIC_L2, _capi.FNB_DTYPE_ANY, None, 0, C.byref(out))
    assert rc in (_capi.FNB_ERR_FORMAT, _capi.FNB_ERR_UNSUPPORTED), (rc, _capi.last_error())
    assert not out.value


def test_array_address_helper_matches_numpy():
    """index._ptr (buffer-protocol fast path of the search wrapper) gives the address ndarray.ctypes.data gives, for the
    shapes the wrapper passes: a batch, the [None, :] view of one query, read-only input, an empty batch."""
    from flatnav_b200.index import _ptr
    a = np.arange(24, dtype=np.float32).reshape(6, 4)
    assert _ptr(a) == a.ctypes.data
    row = a[3][None, :]
    assert row.flags.c_contiguous and _ptr(row) == row.ctypes.data == a.ctypes.data + 3 * 16
    ro = a.copy()
    ro.flags.writeable = False
    assert _ptr(ro) == ro.ctypes.data
    empty = np.empty((0, 4), dtype=np.float32)
    assert _ptr(empty) == empty.ctypes.data
    u8 = np.zeros((2, 32), dtype=np.uint8)
    assert _ptr(u8) == u8.ctypes.data
