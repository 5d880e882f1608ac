"""GPU (-m gpu): the other BASELINE.json configs as parity cases at reduced N (shape, dtype, metric, K and ef of
cfg2 .. cfg5 kept; the graph is built by the unmodified reference on this host): bit-exact against the oracle's
sorted-list twin on a sample, 1e-5 / exact-pair parity and recall within 0.002 against the live reference, run-to-run
determinism, and independence from the lanes-per-row choice for integer data."""
import os

import numpy as np
import pytest

import flatnav_b200
from conftest import build_ref_index, recall, rel_err
from flatnav_b200 import synthetic
from oracle import port, refbin

pytestmark = pytest.mark.gpu

CONFIGS = [
    # id, metric, gen, dim, n, K, efs
    ("cfg2-glove100-ip", "ip", "latent-norm", 100, 120_000, 10, (16, 64, 256, 512)),
    ("cfg3-deep96-l2", "l2", "latent", 96, 150_000, 10, (64, 100)),
    ("cfg4-gist960-l2-k100", "l2", "latent", 960, 30_000, 100, (100, 300)),
    ("cfg5-bigann-u8-l2", "l2", "latent-u8", 128, 200_000, 10, (32, 100, 200)),
]
CLS = {("l2", "float32"): "IndexL2Float", ("ip", "float32"): "IndexIPFloat", ("l2", "uint8"): "IndexL2Uint8"}


@pytest.mark.parametrize("name,metric,gen,dim,n,K,efs", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_config_shaped_parity(ref_cache, name, metric, gen, dim, n, K, efs):
    if (os.cpu_count() or 1) < 8:
        pytest.skip("building the config-shaped graphs needs a few cores")
    path = build_ref_index(ref_cache, metric, gen, n, dim, 32, 100)
    q = synthetic.make(gen, 2000, dim, queries=True)
    ix = getattr(flatnav_b200.index, CLS[(metric, q.dtype.name)]).load_index(path)
    ora = port.OracleIndex(path, port.L2 if metric == "l2" else port.IP)
    _, gt = ix.bruteforce(q, K)
    sample = np.arange(0, 2000, 10)
    cores = os.cpu_count() or 1
    for ef in efs:
        d, l = ix.search(q, K, ef)
        d2, l2 = ix.search(q, K, ef)
        np.testing.assert_array_equal(d, d2)                       # run-to-run determinism
        np.testing.assert_array_equal(l, l2)
        do, lo = ora.search(q[sample], K, ef, mode=port.MODE_LIST, threads=cores)
        np.testing.assert_array_equal(d[sample].view(np.uint32), do.view(np.uint32))
        np.testing.assert_array_equal(l[sample], lo)
        dr, lr, _ = refbin.search(path, metric, q, K, ef, threads=cores)
        if q.dtype == np.float32:
            assert rel_err(d, dr) <= 1e-5
            assert (l == lr).mean() >= 0.999
        else:
            assert (d == dr).mean() >= 0.995                       # integer distances: exact except at ties
        assert abs(recall(l, gt) - recall(lr, gt)) <= 0.002


def test_integer_rows_lane_choice_is_invisible(ref_cache):
    """uint8 rows of <= 128 B are read by 4 lanes instead of 8 (csrc/fnb_layout.h); sums are exact integers, so
    nothing observable may change: FNB_NO_G4 selects the 8-lane kernels for comparison"""
    if (os.cpu_count() or 1) < 8:
        pytest.skip("needs a few cores")
    path = build_ref_index(ref_cache, "l2", "latent-u8", 200_000, 128, 32, 100)
    q = synthetic.make("latent-u8", 3000, 128, queries=True)
    a = flatnav_b200.index.IndexL2Uint8.load_index(path)
    assert a.info["lanes_per_row"] == 4
    os.environ["FNB_NO_G4"] = "1"
    try:
        b = flatnav_b200.index.IndexL2Uint8.load_index(path)
    finally:
        os.environ.pop("FNB_NO_G4")
    assert b.info["lanes_per_row"] == 8
    for K, ef in ((10, 32), (10, 100), (100, 200)):
        da, la = a.search(q, K, ef)
        db, lb = b.search(q, K, ef)
        np.testing.assert_array_equal(da, db)
        np.testing.assert_array_equal(la, lb)
        assert a.last_stats["n_hops"] == b.last_stats["n_hops"]
    ga = a.bruteforce(q[:500], 10)
    gb = b.bruteforce(q[:500], 10)
    np.testing.assert_array_equal(ga[0], gb[0])
    np.testing.assert_array_equal(ga[1], gb[1])
