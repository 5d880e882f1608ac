"""GPU (-m gpu): graph re-ordering, relabelling and Matrix Market link import through the C ABI (csrc/reorder.cu)
against the golden outputs of the UNMODIFIED reference (tests/golden/reorder.json, made by
tools/make_golden_reorder.py): the file saved after `reorder([...])` / `allocate_nodes().build_graph_links()` must be
the reference's byte for byte, and the search on the re-ordered index must stay bit-exact against the oracle."""
import hashlib
import json
import os

import numpy as np
import pytest

import flatnav_b200
from conftest import GOLDEN, build_ref_index, golden_arrays, golden_cases, golden_index_path
from flatnav_b200 import synthetic
from flatnav_b200.data_type import DataType
from oracle import port, refbin

pytestmark = pytest.mark.gpu

CASES = golden_cases()
REORDER = json.load(open(os.path.join(GOLDEN, "reorder.json")))
PM = {"l2": port.L2, "ip": port.IP}
DT = {"f32": DataType.float32, "u8": DataType.uint8, "i8": DataType.int8}


def gpu_class(case):
    return flatnav_b200.index.index_class("l2" if case["metric"] == "l2" else "angular", DT[case["dtype"]])


def sha(path, live_only=False):
    raw = open(path, "rb").read()
    if live_only:  # header + the cur_num_nodes live nodes (the tail is uninitialised memory in the reference)
        raw = raw[:60 + int.from_bytes(raw[20:28], "little") * int.from_bytes(raw[36:44], "little")]
    return hashlib.sha256(raw).hexdigest()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_reordered_file_is_the_references(case, tmp_path):
    g = golden_arrays(case["name"])
    for seq, want in REORDER["reorder"][case["name"]].items():
        ix = gpu_class(case).load_index(golden_index_path(case["name"]))
        ix.reorder(seq.split(","))
        out = str(tmp_path / "r.idx")
        ix.save(out)
        assert sha(out, live_only=True) == want["sha256_live"], seq
        if case["N"] == int.from_bytes(open(out, "rb").read()[28:36], "little"):  # full index: the whole file
            assert sha(out) == want["sha256"], seq
        # the search hot path on the re-ordered index: bit-exact against the oracle on the same file
        ora = port.OracleIndex(out, PM[case["metric"]])
        K, ef = case["runs"][0]
        d, l = ix.search(g["queries"], K, ef)
        do, lo = ora.search(g["queries"], K, ef, mode=port.MODE_LIST)
        np.testing.assert_array_equal(d.view(np.uint32), do.view(np.uint32))
        np.testing.assert_array_equal(l, lo)


def test_outdegree_table_and_relabel_round_trip(tmp_path):
    case = CASES[0]
    path = golden_index_path(case["name"])
    ix = gpu_class(case).load_index(path)
    links = port.OracleIndex(path, PM[case["metric"]]).links()
    table = ix.get_graph_outdegree_table()  # Index.h:240-251: self-loops dropped, slot order kept
    assert len(table) == links.shape[0]
    for n in (0, 1, 17, links.shape[0] - 1):
        assert table[n] == [int(x) for x in links[n] if x != n]
    rng = np.random.default_rng(3)
    perm = rng.permutation(links.shape[0]).astype(np.uint32)
    g = golden_arrays(case["name"])
    K, ef = 10, links.shape[0]  # exhaustive beam: the result set no longer depends on the entry point
    d0, l0 = ix.search(g["queries"], K, ef)
    ix.relabel(perm)
    d1, l1 = ix.search(g["queries"], K, ef)
    np.testing.assert_array_equal(d0, d1)  # labels travel with their nodes; equal distances may swap places
    assert (l0 == l1).mean() > 0.99
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size, dtype=np.uint32)
    ix.relabel(inv)
    out = str(tmp_path / "back.idx")
    ix.save(out)
    assert open(out, "rb").read() == open(path, "rb").read()
    bad = perm.copy()
    bad[0] = bad[1]
    with pytest.raises(ValueError, match="not a permutation"):
        ix.relabel(bad)
    with pytest.raises(ValueError):
        ix.relabel(perm[:-1])


def test_reorder_validation_matches_the_binding():
    case = CASES[0]
    ix = gpu_class(case).load_index(golden_index_path(case["name"]))
    with pytest.raises(ValueError, match="`bogus` is not a supported graph re-ordering strategy."):
        ix.reorder(["gorder", "bogus"])  # bindings.cpp:287-293: validated before anything is applied
    with pytest.raises(ValueError, match="Invalid reordering method: GORDER"):
        ix.reorder(["GORDER"])  # passes the case-insensitive check, fails Index.h:417-423
    ix.reorder([])


def test_mtx_import_is_the_references(tmp_path):
    m = REORDER["mtx"]
    data = np.load(os.path.join(GOLDEN, "mtx_case.npz"))["data"]
    mtx = os.path.join(GOLDEN, "mtx_case.mtx")
    ix = flatnav_b200.index.create("l2", m["D"], m["N"], m["M"])
    assert ix.allocate_nodes(data) is ix  # chaining, run-benchmark.py:236
    ix.build_graph_links(mtx)
    out = str(tmp_path / "m.idx")
    ix.save(out)
    assert sha(out) == m["sha256"]
    d, l = ix.search(data[:5], 1, 16)
    assert d.shape == (5, 1)
    # error behaviour of Index::buildGraphLinks (Index.h:187-217)
    with pytest.raises(RuntimeError, match="Unable to open file for reading"):
        ix.build_graph_links(str(tmp_path / "missing.mtx"))
    other = flatnav_b200.index.create("l2", m["D"], m["N"] + 1, m["M"])
    with pytest.raises(RuntimeError, match="Number of vertices in the mtx file does not match"):
        other.build_graph_links(mtx)
    other = flatnav_b200.index.create("l2", m["D"], m["N"], m["M"] + 1)
    with pytest.raises(RuntimeError, match="Number of edges in the mtx file does not match"):
        other.build_graph_links(mtx)
    with pytest.raises(ValueError, match="Data has incorrect dimensions."):
        ix.allocate_nodes(data[:, :-1])
    with pytest.raises(ValueError, match="Maximum number of nodes reached"):
        ix.allocate_nodes(data[:1])


@pytest.mark.skipif(not refbin.available(), reason="oracle/_ref reference binary not available")
def test_reorder_live_reference_20k(ref_cache, tmp_path):
    """a multi-threaded reference build at M=32, re-ordered by the reference and by this engine"""
    path = build_ref_index(ref_cache, "l2", "latent", 20000, 96, 32, 100)
    q = synthetic.make("latent", 200, 96, queries=True)
    for seq in (["gorder"], ["rcm"]):
        want = str(tmp_path / "ref.idx")
        refbin.reorder(path, "l2", "f32", seq, want)
        ix = flatnav_b200.index.IndexL2Float.load_index(path)
        ix.reorder(seq)
        got = str(tmp_path / "got.idx")
        ix.save(got)
        assert open(got, "rb").read() == open(want, "rb").read(), seq
        dr, lr, _ = refbin.search(want, "l2", q, 10, 64, threads=1)
        d, l = ix.search(q, 10, 64)
        assert float(np.max(np.abs(d - dr) / np.maximum(np.abs(dr), 1e-6))) <= 1e-5
        assert (l == lr).mean() >= 0.999
