"""GPU (-m gpu): the tcgen05 exact-scan path (csrc/bruteforce_tc.cu: tensor-core filter + exact re-rank) must
return, bit for bit, what the CUDA-core exact scan and the CPU oracle scan return: top-K by (distance, node id),
distances in the traversal kernel's arithmetic (BASELINE.json: "brute-force ground truth and top-k IDs are
bit-exact against a CPU exact scan")."""
import os

import numpy as np
import pytest

import flatnav_b200
from flatnav_b200 import synthetic
from oracle import port
from tools.rawindex import index_bytes

pytestmark = pytest.mark.gpu

CLS = {("l2", np.float32): "IndexL2Float", ("ip", np.float32): "IndexIPFloat", ("l2", np.uint8): "IndexL2Uint8",
       ("ip", np.uint8): "IndexIPUint8", ("l2", np.int8): "IndexL2Int8", ("ip", np.int8): "IndexIPInt8"}


def run_mode(ix, q, K, mode, splits=None):
    old = {k: os.environ.get(k) for k in ("FNB_BF_MODE", "FNB_BF_SPLITS")}
    os.environ["FNB_BF_MODE"] = mode
    if splits:
        os.environ["FNB_BF_SPLITS"] = str(splits)
    try:
        d, l = ix.bruteforce(q, K)
        return d, l, dict(ix.last_bruteforce_stats)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def make_index(metric, data, labels=None):
    blob = index_bytes(data, M=4, labels=labels)
    cls = getattr(flatnav_b200.index, CLS[(metric, data.dtype.type)])
    return cls.from_bytes(blob), blob


CASES = [
    # metric, generator, N, D, Q, K, max unsafe fraction
    ("l2", "latent", 50_000, 128, 300, 10, 0.05),
    ("ip", "latent-norm", 30_001, 100, 257, 10, 0.05),
    ("l2", "latent", 20_000, 960, 130, 100, 0.10),
    ("l2", "iid", 5_000, 7, 64, 5, 0.20),
    ("l2", "iid", 33_000, 24, 1, 10, 1.0),
    ("l2", "latent-u8", 40_000, 128, 200, 10, 0.05),
    ("ip", "latent-u8", 9_999, 32, 129, 10, 1.0),
    ("l2", "latent-i8", 20_000, 128, 128, 10, 0.05),
    ("ip", "latent-i8", 12_345, 48, 100, 20, 1.0),
    ("l2", "latent-u8", 6_000, 960, 40, 10, 1.0),
]


@pytest.mark.parametrize("metric,gen,n,d,nq,K,max_unsafe", CASES)
def test_tensor_path_bit_exact(metric, gen, n, d, nq, K, max_unsafe):
    data = synthetic.make(gen, n, d)
    q = synthetic.make(gen, nq, d, queries=True)
    labels = (np.arange(n, dtype=np.int32) * 7 + 3)               # labels come from the label field
    ix, blob = make_index(metric, data, labels)
    dt, lt, st = run_mode(ix, q, K, "tensor")
    assert st["path"] == 1 and st["gemm_flops"] > 0
    de, le, se = run_mode(ix, q, K, "exact")
    assert se["path"] == 0
    np.testing.assert_array_equal(dt.view(np.uint32), de.view(np.uint32))
    np.testing.assert_array_equal(lt, le)
    assert st["n_unsafe"] <= max_unsafe * nq, st
    ora = port.OracleIndex(blob, port.L2 if metric == "l2" else port.IP)
    sel = slice(0, min(nq, 16))
    do, lo = ora.bruteforce(q[sel], K)
    np.testing.assert_array_equal(dt[sel].view(np.uint32), do.view(np.uint32))
    np.testing.assert_array_equal(lt[sel], lo)


def test_tensor_path_many_slices_and_ties():
    """duplicated rows (exact ties at the K-th distance) and every true neighbour inside ONE slice: the margin
    proof must send such queries to the exact re-scan instead of returning a wrong list"""
    rng = np.random.default_rng(5)
    base = synthetic.make("latent", 20_000, 64)
    data = np.concatenate([base, base[:3000], base[:3000]])       # triplicates
    q = base[:200] + 1e-3 * rng.standard_normal((200, 64)).astype(np.float32)
    ix, blob = make_index("l2", data)
    for splits in (1, 7, 40):
        dt, lt, st = run_mode(ix, q, 10, "tensor", splits)
        de, le, _ = run_mode(ix, q, 10, "exact")
        np.testing.assert_array_equal(dt.view(np.uint32), de.view(np.uint32))
        np.testing.assert_array_equal(lt, le)
    # clustered order: sort rows by distance to q[0] so its neighbours are contiguous
    order = np.argsort(((base - q[0]) ** 2).sum(1))
    ix2, _ = make_index("l2", np.ascontiguousarray(base[order]))
    dt, lt, st = run_mode(ix2, q, 50, "tensor", 16)
    de, le, _ = run_mode(ix2, q, 50, "exact")
    np.testing.assert_array_equal(dt.view(np.uint32), de.view(np.uint32))
    np.testing.assert_array_equal(lt, le)


def test_default_dispatch_uses_tensor_path_when_large():
    data = synthetic.make("latent", 200_000, 128)
    q = synthetic.make("latent", 1000, 128, queries=True)
    ix, _ = make_index("l2", data)
    d, l = ix.bruteforce(q, 10)
    st = ix.last_bruteforce_stats
    assert st["path"] == 1 and st["n_unsafe"] <= 50
    gt = np.argsort(((data[None, :, :] - q[:4, None, :]) ** 2).sum(-1), axis=1)[:, :10]
    assert (l[:4] == gt).mean() >= 0.95                           # float64-free sanity: same neighbours
    d2, l2 = ix.bruteforce(q[:8], 10)                              # small problem: CUDA-core scan
    assert ix.last_bruteforce_stats["path"] == 0
    np.testing.assert_array_equal(d2, d[:8])
    np.testing.assert_array_equal(l2, l[:8])
