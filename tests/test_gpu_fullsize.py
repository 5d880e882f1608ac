"""GPU (-m gpu): parity at the FULL size of BASELINE.json configs[0] (1M x 128 float32, M=32, 10k queries, K=10,
ef=100) through properties that do not need a full-size CPU run, plus a sampled comparison against the oracle
and the live reference.  The index is built once on this host by the unmodified reference (about 12 s on 16
cores) — construction is out of scope of the product."""
import os

import numpy as np
import pytest

import flatnav_b200
from conftest import recall, rel_err
from flatnav_b200 import synthetic
from oracle import port, refbin

pytestmark = pytest.mark.gpu

N, D, M, EFC, Q, K, EF = 1_000_000, 128, 32, 100, 10_000, 10, 100


@pytest.fixture(scope="module")
def full():
    if not refbin.available():
        pytest.skip("reference builder (oracle/_ref) not runnable on this host")
    if (os.cpu_count() or 1) < 8:
        pytest.skip("full-size build needs a few cores")
    from tools.workload import ensure_index
    path, _ = ensure_index("latent", N, D, "l2", M, EFC)
    ix = flatnav_b200.index.IndexL2Float.load_index(path)
    q = synthetic.make("latent", Q, D, queries=True)
    d, l = ix.search(q, K, EF)
    return dict(path=path, ix=ix, q=q, d=d, l=l, stats=dict(ix.last_stats))


def test_fullsize_structure(full):
    d, l = full["d"], full["l"]
    assert d.shape == (Q, K) and l.shape == (Q, K)
    assert np.all(np.diff(d, axis=1) >= 0)                      # ascending (Index.h:402-406)
    assert l.min() >= 0 and l.max() < N
    assert all(len(set(r.tolist())) == K for r in l[::97])      # no duplicate labels in a row
    st = full["stats"]
    assert st["n_short"] == 0 and st["n_queries"] == Q
    assert 95 <= st["n_hops"] / Q <= 115                        # hops ~ ef + 3
    assert 1900 <= st["n_dist"] / Q <= 2300                     # ~2073 incl. the 100 entry probes


def test_fullsize_pairs_are_exact(full):
    """every returned distance is the squared L2 distance of the node that carries the returned label.  The
    multi-threaded addBatch of the reference hands out node ids in lock order (Index.h:262-271, 364), so label !=
    node id for a few nodes: map labels back to nodes through the label field."""
    ora = port.OracleIndex(full["path"], port.L2)
    vec, lab = ora.vectors(), ora.labels()
    node_of = np.empty(N, dtype=np.int64)
    node_of[lab] = np.arange(N)
    assert np.array_equal(np.sort(lab), np.arange(N))            # labels are a permutation of the row numbers
    for i in range(0, Q, 50):
        exact = np.sum((vec[node_of[full["l"][i]]].astype(np.float64) - full["q"][i].astype(np.float64)) ** 2, axis=1)
        assert rel_err(full["d"][i], exact) <= 1e-5


def test_fullsize_determinism_and_batch_independence(full):
    ix, q = full["ix"], full["q"]
    d2, l2 = ix.search(q, K, EF)
    np.testing.assert_array_equal(d2, full["d"])               # idempotent
    np.testing.assert_array_equal(l2, full["l"])
    perm = np.random.default_rng(0).permutation(Q)
    dp, lp = ix.search(q[perm], K, EF)                           # a query's result does not depend on its batch slot
    np.testing.assert_array_equal(dp, full["d"][perm])
    np.testing.assert_array_equal(lp, full["l"][perm])
    for i in (0, 1234, Q - 1):                                   # search_single == row of search
        d1, l1 = ix.search_single(q[i], K, EF)
        np.testing.assert_array_equal(d1, full["d"][i])
        np.testing.assert_array_equal(l1, full["l"][i])
    d100, l100 = ix.search(q[:2000], 100, EF)                    # buffer = max(ef, K): K=10 is a prefix of K=100
    np.testing.assert_array_equal(d100[:, :K], full["d"][:2000])
    np.testing.assert_array_equal(l100[:, :K], full["l"][:2000])


def test_fullsize_sample_vs_oracle_and_reference(full):
    ix, q = full["ix"], full["q"]
    sample = np.arange(0, Q, 40)                                 # 250 queries
    ora = port.OracleIndex(full["path"], port.L2)
    do, lo, nd, nh = ora.search(q[sample], K, EF, mode=port.MODE_LIST, counters=True, threads=os.cpu_count() or 1)
    np.testing.assert_array_equal(full["d"][sample].view(np.uint32), do.view(np.uint32))   # bit-exact vs the oracle twin
    np.testing.assert_array_equal(full["l"][sample], lo)
    dr, lr, _ = refbin.search(full["path"], "l2", q[sample], K, EF, threads=os.cpu_count() or 1)
    assert rel_err(full["d"][sample], dr) <= 1e-5               # BASELINE.json: 1e-5 relative vs the reference
    assert (full["l"][sample] == lr).mean() >= 0.999


def test_fullsize_recall_against_bruteforce(full):
    ix, q = full["ix"], full["q"]
    sample = np.arange(0, Q, 10)                                 # 1000 queries of exact ground truth on the GPU
    gd, gl = ix.bruteforce(q[sample], K)
    ora = port.OracleIndex(full["path"], port.L2)
    od, ol = ora.bruteforce(q[sample[:20]], K)                   # the GPU scan itself is bit-exact vs the CPU scan
    np.testing.assert_array_equal(gd[:20].view(np.uint32), od.view(np.uint32))
    np.testing.assert_array_equal(gl[:20], ol)
    r100 = recall(full["l"][sample], gl)
    assert r100 >= 0.95                                          # the operating point of the metric
    prev = 0.0
    for ef in (16, 32, 64, 100, 200):                            # recall grows with ef; reference recall matches
        _, l = ix.search(q[sample], K, ef)
        r = recall(l, gl)
        assert r >= prev - 0.002
        prev = r
        _, lr, _ = refbin.search(full["path"], "l2", q[sample], K, ef, threads=os.cpu_count() or 1)
        assert abs(r - recall(lr, gl)) <= 0.002                  # BASELINE.json: recall within 0.002 at every ef
