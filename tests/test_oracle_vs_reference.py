"""CPU: the oracle port against the UNMODIFIED reference run live (oracle/_ref, compiled from the reference's
own headers by oracle/Makefile) on freshly built indexes — all six metric x dtype instantiations that
include/flatnav/tests/test_serialization.cpp:78-176 covers."""
import numpy as np
import pytest

from conftest import build_ref_index, recall, rel_err
from flatnav_b200 import synthetic
from oracle import port, refbin

pytestmark = pytest.mark.skipif(not refbin.available(), reason="oracle/_ref reference binary not available")

COMBOS = [
    ("l2", "latent", 128, port.L2), ("ip", "latent-norm", 100, port.IP),
    ("l2", "latent-u8", 128, port.L2), ("ip", "latent-u8", 64, port.IP),
    ("l2", "latent-i8", 64, port.L2), ("ip", "latent-i8", 64, port.IP),
]


@pytest.mark.parametrize("metric,gen,dim,pm", COMBOS, ids=[f"{m}-{g}-{d}" for m, g, d, _ in COMBOS])
def test_oracle_vs_live_reference(ref_cache, metric, gen, dim, pm):
    n, M, efc, nq, K = 6000, 32, 100, 200, 10
    path = build_ref_index(ref_cache, metric, gen, n, dim, M, efc)
    queries = synthetic.make(gen, nq, dim, queries=True)
    ix = port.OracleIndex(path, pm)
    for ef in (16, 100):
        dr, lr, _ = refbin.search(path, metric, queries, K, ef, threads=1)
        d, l = ix.search(queries, K, ef, mode=port.MODE_LIST)
        if queries.dtype == np.float32:
            assert rel_err(d, dr) <= 1e-5
            assert (l == lr).mean() >= 0.999
        else:
            np.testing.assert_array_equal(d, dr)
            diff = l != lr
            assert np.all(d[diff] == dr[diff])
        gt_d, gt_l = ix.bruteforce(queries, K)
        assert abs(recall(l, gt_l) - recall(lr, gt_l)) <= 0.002  # BASELINE.json recall criterion


def test_reference_multithreaded_loop_equals_single(ref_cache):
    """executeInParallel fan-out (bindings.cpp:198-211) returns what the serial loop returns"""
    path = build_ref_index(ref_cache, "l2", "latent", 6000, 128, 32, 100)
    q = synthetic.make("latent", 200, 128, queries=True)
    d1, l1, _ = refbin.search(path, "l2", q, 10, 64, threads=1)
    d4, l4, info = refbin.search(path, "l2", q, 10, 64, threads=4)
    np.testing.assert_array_equal(d1, d4)
    np.testing.assert_array_equal(l1, l4)
    assert info["short_results"] == 0
