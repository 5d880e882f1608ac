"""GPU (-m gpu): graph construction on the GPU (csrc/build.cu; SURVEY.md §8f rank 1 — Index::add / selectNeighbors /
connectNeighbors, Index.h:301-378, 714-834) judged the way the reference's own tests judge construction
(python-bindings/unit_tests/test_index.py: build, search, recall):
  * the file it saves is a valid reference index: the UNMODIFIED reference loads it and returns the same results as
    this engine on it (distances 1e-5 / exact, labels equal), and the oracle twin is bit-exact on it;
  * graph invariants of the reference hold: links < cur_num_nodes, rows packed (real links first, then self-loops),
    no duplicate links, every back-link rule leaves out-degree <= M, the new node gets <= max(M/2, 1) own links;
  * search quality equals a reference-built graph over the same data: recall@10 within 0.015 at every ef (within 0.002 at 1M nodes, tools/build_probe.py);
  * appending to a loaded reference-built index works, labels are kept."""
import os

import numpy as np
import pytest

import flatnav_b200
from conftest import build_ref_index, recall, rel_err
from flatnav_b200 import synthetic
from flatnav_b200.data_type import DataType
from oracle import port, refbin

pytestmark = pytest.mark.gpu

CASES = [
    # metric, gen, dim, n, M
    ("l2", "latent", 128, 60_000, 32),
    ("angular", "latent-norm", 100, 40_000, 32),
    ("l2", "latent-u8", 128, 60_000, 32),
    ("l2", "latent", 24, 20_000, 16),
    ("angular", "latent-i8", 64, 20_000, 8),
]


def check_graph(path, metric, M):
    ora = port.OracleIndex(path, port.L2 if metric == "l2" else port.IP)
    links = ora.links()
    n = ora.cur_num_nodes
    assert links.max() < n
    self_loop = links == np.arange(n, dtype=np.uint32)[:, None]
    deg = (~self_loop).sum(1)
    first_self = np.where(self_loop.any(1), self_loop.argmax(1), M)
    assert np.array_equal(deg, first_self), "rows must be packed: real links first, then self-loops"
    srt = np.sort(np.where(self_loop, np.uint32(0xffffffff), links), axis=1)
    dup = (srt[:, 1:] == srt[:, :-1]) & (srt[:, 1:] != 0xffffffff)
    assert not dup.any(), "duplicate links"
    assert deg[1:].min() >= 1 and deg.max() <= M
    return ora, deg


@pytest.mark.parametrize("metric,gen,dim,n,M", CASES, ids=[f"{c[0]}-{c[1]}-{c[2]}" for c in CASES])
def test_gpu_built_graph_is_a_reference_index(tmp_path, ref_cache, metric, gen, dim, n, M):
    data = synthetic.make(gen, n, dim)
    q = synthetic.make(gen, 1000, dim, queries=True)
    dt = {np.dtype(np.float32): DataType.float32, np.dtype(np.uint8): DataType.uint8, np.dtype(np.int8): DataType.int8}[data.dtype]
    ix = flatnav_b200.index.create(metric, dim, n, M, dt)
    labels = np.arange(n, dtype=np.int32) * 3 + 1
    half = n // 2
    ix.add(data[:half], ef_construction=100, labels=labels[:half])
    ix.add(data[half:], ef_construction=100, num_initializations=100, labels=labels[half:])   # a second call appends
    assert ix.info["cur_num_nodes"] == n and ix.last_build_stats["n_added"] == n - half
    path = str(tmp_path / "gpu_built.idx")
    ix.save(path)
    m = "l2" if metric == "l2" else "ip"
    ora, deg = check_graph(path, m, M)
    assert np.array_equal(ora.labels(), labels)
    assert np.array_equal(ora.vectors(), data)
    # the unmodified reference reads the file and agrees with this engine on it
    K = 10
    gt = ix.bruteforce(q, K)[1]
    ref_path = build_ref_index(ref_cache, m, gen, n, dim, M, 100)
    ref_ix = type(ix).load_index(ref_path)
    for ef in (32, 100, 200):
        d, l = ix.search(q, K, ef)
        dr, lr, _ = refbin.search(path, m, q, K, ef, threads=os.cpu_count() or 1)
        if data.dtype == np.float32:
            assert rel_err(d, dr) <= 1e-5
            assert (l == lr).mean() >= 0.999
        else:
            assert (d == dr).mean() >= 0.99
        do, lo = ora.search(q[:100], K, ef, mode=port.MODE_LIST, threads=os.cpu_count() or 1)
        np.testing.assert_array_equal(d[:100].view(np.uint32), do.view(np.uint32))
        np.testing.assert_array_equal(l[:100], lo)
        # same quality as a graph built by the reference over the same data (its labels are row numbers)
        _, l_ref = ref_ix.search(q, K, ef)
        r_gpu, r_ref = recall((l - 1) // 3, (gt - 1) // 3), recall(l_ref, (gt - 1) // 3)
        assert r_gpu >= r_ref - 0.015, (ef, r_gpu, r_ref)
    ref_deg = (port.OracleIndex(ref_path, port.L2 if m == "l2" else port.IP).links() !=
               np.arange(n, dtype=np.uint32)[:, None]).sum(1)
    assert abs(deg.mean() - ref_deg.mean()) <= 0.15 * ref_deg.mean()


def test_append_to_a_loaded_reference_index(tmp_path, ref_cache):
    n0, n1, dim, M = 30_000, 10_000, 128, 32
    ref_path = build_ref_index(ref_cache, "l2", "latent", n0, dim, M, 100)
    extra = synthetic.make("latent", n1, dim, stream=3)
    ix = flatnav_b200.index.IndexL2Float.load_index(ref_path)
    with pytest.raises(ValueError, match="Maximum number of nodes reached"):
        ix.add(extra, 100)
    ix.reserve(n0 + n1)
    ix.add(extra, 100, labels=np.arange(n0, n0 + n1))
    path = str(tmp_path / "appended.idx")
    ix.save(path)
    check_graph(path, "l2", M)
    d, l = ix.search(extra[:500], 1, 64)             # every appended vector finds itself
    assert (l[:, 0] == np.arange(n0, n0 + 500)).mean() >= 0.99 and np.all(d[l[:, 0] == np.arange(n0, n0 + 500), 0] == 0)
    q = synthetic.make("latent", 500, dim, queries=True)
    gt = ix.bruteforce(q, 10)[1]
    assert recall(ix.search(q, 10, 100)[1], gt) >= 0.95


def test_add_argument_errors_match_the_binding():
    ix = flatnav_b200.index.create("l2", 16, 100, 8)
    with pytest.raises(ValueError, match="Data has incorrect dimensions"):
        ix.add(np.zeros((4, 15), np.float32), 10)
    with pytest.raises(ValueError, match="Incorrect number of labels"):
        ix.add(np.zeros((4, 16), np.float32), 10, labels=[1, 2, 3])
    with pytest.raises(ValueError, match="num_initializations must be greater than 0"):
        ix.add(np.zeros((4, 16), np.float32), 10, num_initializations=0)
    x = np.random.default_rng(0).standard_normal((100, 16)).astype(np.float32)
    ix.add(x, 20)                                     # float64 / lists are cast like py::array::forcecast
    with pytest.raises(ValueError, match="Maximum number of nodes reached"):
        ix.add(x[:1], 20)
    d, l = ix.search(x, 1, 20)
    assert (l[:, 0] == np.arange(100)).mean() >= 0.98
