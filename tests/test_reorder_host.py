"""CPU: graph re-ordering and Matrix Market link import.

* the oracle's literal restatement (oracle/flatnav_oracle.cpp: ora_reorder, ora_build_graph_links) against the golden
  outputs of the UNMODIFIED reference (tests/golden/reorder.json, tools/make_golden_reorder.py) and, where oracle/_ref
  can run, against the reference live on a fresh index: files byte-identical;
* the product's host-side ordering step (fnb_graph_order: O(1) boundary-table queue, CSR tables) against the oracle's
  sorted-vector restatement: identical permutations, also on graphs with duplicate links, sparse rows and isolated
  nodes.  (The device half — relabel kernel, save — is tests/test_gpu_reorder.py.)
"""
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

from conftest import GOLDEN, build_ref_index, golden_cases, golden_index_path
from flatnav_b200 import _capi
from oracle import port, refbin

REORDER = json.load(open(os.path.join(GOLDEN, "reorder.json")))
SEQS = sorted(next(iter(REORDER["reorder"].values())).keys())


def product_order(links: np.ndarray, method: int, window: int = 5) -> np.ndarray:
    links = np.ascontiguousarray(links, dtype=np.uint32)
    perm = np.empty(links.shape[0], dtype=np.uint32)
    rc = _capi.lib().fnb_graph_order(links.ctypes.data, links.shape[0], links.shape[1], method, window,
                                     perm.ctypes.data)
    assert rc == 0, _capi.last_error()
    return perm


def image_from_links(links: np.ndarray) -> bytes:
    """A minimal index file (4-byte vectors) around a link table, for the oracle."""
    n, M = links.shape
    ds, ns = 4, 4 + 4 * M + 4
    nodes = np.zeros((n, ns), dtype=np.uint8)
    nodes[:, 4:4 + 4 * M] = np.ascontiguousarray(links, dtype=np.uint32).view(np.uint8).reshape(n, 4 * M)
    nodes[:, 4 + 4 * M:] = np.arange(n, dtype=np.int32).view(np.uint8).reshape(n, 4)
    hdr = np.array([9], dtype=np.int32).tobytes() + np.array([M, ds, ns, n, n, 1, ds], dtype=np.uint64).tobytes()
    return hdr + nodes.tobytes()


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_oracle_reorder_equals_reference_golden(case):
    blob = open(golden_index_path(case["name"]), "rb").read()
    for seq in SEQS:
        out, perms = port.reorder_file(blob, seq.split(","))
        want = REORDER["reorder"][case["name"]][seq]
        assert hashlib.sha256(out.tobytes()).hexdigest() == want["sha256"], seq
        total = perms[0]
        for p in perms[1:]:
            total = p[total]  # node i sits at total[i] after the earlier steps and moves to p[total[i]]
        assert zlib.crc32(total.tobytes()) == want["perm_crc32"]


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_product_order_equals_oracle_on_golden(case):
    blob = open(golden_index_path(case["name"]), "rb").read()
    links = port.OracleIndex(blob, port.L2).links()
    for method, name in ((_capi.FNB_REORDER_GORDER, "gorder"), (_capi.FNB_REORDER_RCM, "rcm")):
        _, perms = port.reorder_file(blob, [name])
        np.testing.assert_array_equal(product_order(links, method), perms[0])
        assert zlib.crc32(perms[0].tobytes()) == REORDER["reorder"][case["name"]][name]["perm_crc32"]


@pytest.mark.parametrize("n,M,fill,seed", [(1, 4, 1.0, 0), (2, 3, 1.0, 1), (97, 5, 0.6, 2), (400, 8, 0.3, 3),
                                            (1500, 16, 0.9, 4), (3000, 32, 0.5, 5)])
def test_product_order_equals_oracle_on_irregular_graphs(n, M, fill, seed):
    """duplicate links, rows with gaps (self-loops between real links), isolated nodes, hubs"""
    rng = np.random.default_rng(seed)
    hubs = rng.integers(0, n, max(1, n // 50))
    links = np.where(rng.random((n, M)) < 0.3, hubs[rng.integers(0, hubs.size, (n, M))], rng.integers(0, n, (n, M)))
    own = np.arange(n)[:, None]
    links = np.where(rng.random((n, M)) < fill, links, own).astype(np.uint32)
    links[rng.integers(0, n, max(1, n // 20))] = own[:1] * 0 + np.arange(n)[rng.integers(0, n, max(1, n // 20))][:, None]
    links[::11] = own[::11]  # isolated: every slot a self-loop
    blob = image_from_links(links)
    for window in (1, 5, 9):
        _, perms = port.reorder_file(blob, ["gorder"], window=window)
        np.testing.assert_array_equal(product_order(links, _capi.FNB_REORDER_GORDER, window), perms[0])
    _, perms = port.reorder_file(blob, ["rcm"])
    np.testing.assert_array_equal(product_order(links, _capi.FNB_REORDER_RCM), perms[0])


def test_graph_order_rejects_bad_input():
    links = np.array([[1, 5], [0, 1]], dtype=np.uint32)
    perm = np.empty(2, dtype=np.uint32)
    lib = _capi.lib()
    assert lib.fnb_graph_order(links.ctypes.data, 2, 2, 0, 5, perm.ctypes.data) == _capi.FNB_ERR_INVALID_ARG
    assert "outside" in _capi.last_error()
    links[0, 1] = 0
    assert lib.fnb_graph_order(links.ctypes.data, 2, 2, 7, 5, perm.ctypes.data) == _capi.FNB_ERR_INVALID_ARG
    assert _capi.last_error().startswith("Invalid reordering method")  # Index.h:421-423


def test_oracle_mtx_import_equals_reference_golden():
    m = REORDER["mtx"]
    data = np.load(os.path.join(GOLDEN, "mtx_case.npz"))["data"]
    n, d, M = m["N"], m["D"], m["M"]
    ds, ns = 4 * d, 4 * d + 4 * M + 4
    nodes = np.zeros((n, ns), dtype=np.uint8)  # allocateNode: vector, self-loops, label (Index.h:262-272)
    nodes[:, :ds] = data.view(np.uint8).reshape(n, ds)
    nodes[:, ds:ds + 4 * M] = np.repeat(np.arange(n, dtype=np.uint32)[:, None], M, 1).view(np.uint8).reshape(n, 4 * M)
    nodes[:, ds + 4 * M:] = np.arange(n, dtype=np.int32).view(np.uint8).reshape(n, 4)
    hdr = np.array([9], dtype=np.int32).tobytes() + np.array([M, ds, ns, n, n, d, ds], dtype=np.uint64).tobytes()
    rows = [ln.split() for ln in open(os.path.join(GOLDEN, "mtx_case.mtx")) if not ln.startswith("%")][1:]
    src = np.array([int(r[0]) - 1 for r in rows], dtype=np.uint32)
    dst = np.array([int(r[1]) - 1 for r in rows], dtype=np.uint32)
    out = port.build_graph_links_file(hdr + nodes.tobytes(), src, dst)
    assert hashlib.sha256(out.tobytes()).hexdigest() == m["sha256"]


@pytest.mark.skipif(not refbin.available(), reason="oracle/_ref reference binary not available")
@pytest.mark.parametrize("metric,gen,dim,dt", [("l2", "latent", 128, "f32"), ("ip", "latent-u8", 64, "u8")])
def test_oracle_and_product_order_vs_live_reference(ref_cache, tmp_path, metric, gen, dim, dt):
    """a multi-threaded reference build (M=32: neighbour lists longer than std::sort's insertion-sort threshold)"""
    path = build_ref_index(ref_cache, metric, gen, 6000, dim, 32, 100)
    blob = open(path, "rb").read()
    links = port.OracleIndex(blob, port.L2).links()
    for seq in (["gorder"], ["rcm"], ["rcm", "gorder"]):
        out = str(tmp_path / "o.idx")
        refbin.reorder(path, metric, dt, seq, out)
        mine, perms = port.reorder_file(blob, seq)
        assert mine.tobytes() == open(out, "rb").read(), seq
        if len(seq) == 1:
            method = _capi.FNB_REORDER_GORDER if seq[0] == "gorder" else _capi.FNB_REORDER_RCM
            np.testing.assert_array_equal(product_order(links, method), perms[0])
