"""CPU: the oracle port against the golden outputs of the unmodified reference (tests/golden, made by
tools/make_golden.py) and against the distance definitions the reference's own tests pin
(include/flatnav/tests/test_distances.cpp)."""
import numpy as np
import pytest

from conftest import golden_arrays, golden_cases, golden_index_path, rel_err
from oracle import port

CASES = golden_cases()
METRIC = {"l2": port.L2, "ip": port.IP}


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
@pytest.mark.parametrize("mode", [port.MODE_HEAPS, port.MODE_LIST], ids=["heaps", "list"])
def test_oracle_matches_reference_golden(case, mode):
    g = golden_arrays(case["name"])
    ix = port.OracleIndex(golden_index_path(case["name"]), METRIC[case["metric"]])
    assert ix.M == case["M"] and ix.dim == case["D"] and ix.cur_num_nodes == case["N"]
    assert ix.node_size_bytes == ix.data_size_bytes + 4 * ix.M + 4  # test_serialization.cpp:56
    for K, ef in case["runs"]:
        d, l = ix.search(g["queries"], K, ef, mode=mode)
        dr, lr = g[f"dist_k{K}_ef{ef}"], g[f"label_k{K}_ef{ef}"]
        if case["dtype"] == "f32":
            assert rel_err(d, dr) <= 1e-5  # BASELINE.json: distances match the reference to 1e-5 relative
            assert (l == lr).mean() >= 0.999
        else:
            np.testing.assert_array_equal(d, dr)  # integer arithmetic: bit-exact
            # labels may differ only inside groups of exactly tied distances
            diff = l != lr
            assert np.all(d[diff] == dr[diff])
            assert diff.mean() <= 0.02


@pytest.mark.parametrize("case", [c for c in CASES if c["dtype"] == "f32"], ids=lambda c: c["name"])
def test_two_formulations_agree(case):
    """two-heap restatement == sorted-list formulation (what the CUDA kernel implements), incl. counters"""
    g = golden_arrays(case["name"])
    ix = port.OracleIndex(golden_index_path(case["name"]), METRIC[case["metric"]])
    K, ef = case["runs"][0]
    a = ix.search(g["queries"], K, ef, mode=port.MODE_HEAPS, counters=True)
    b = ix.search(g["queries"], K, ef, mode=port.MODE_LIST, counters=True)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


def test_header_layout_of_golden_file():
    """60-byte cereal header: int32 dtype | u64 M, data_size, node_size, max_nodes, cur_nodes | u64 dim, data_size"""
    raw = np.fromfile(golden_index_path("l2_f32_d24"), dtype=np.uint8)
    assert int(raw[:4].view(np.int32)[0]) == 9  # DataType::float32
    v = raw[4:60].view(np.uint64)
    assert list(v) == [16, 96, 96 + 64 + 4, 2000, 2000, 24, 96]
    assert raw.size == 60 + 164 * 2000


@pytest.mark.parametrize("dim", [128, 100, 37, 7, 960])
def test_distance_orders_agree(dim):
    """test_distances.cpp:37-178 checks every SIMD kernel against the scalar definition within 1e-2 on
    N(0,10^2) vectors; the same check for the oracle's two summation orders, at 1e-5 relative."""
    rng = np.random.default_rng(dim)
    for _ in range(20):
        x = (rng.standard_normal(dim) * 10).astype(np.float32)
        y = (rng.standard_normal(dim) * 10).astype(np.float32)
        for metric in (port.L2, port.IP):
            a = port.distance(x, y, metric, port.ORDER_SEQUENTIAL)
            b = port.distance(x, y, metric, port.ORDER_LANES)
            exact = float(np.sum((x.astype(np.float64) - y) ** 2)) if metric == port.L2 else 1.0 - float(
                np.dot(x.astype(np.float64), y))
            scale = float(np.sum(np.abs(x.astype(np.float64) * y))) if metric == port.IP else abs(exact)
            assert abs(a - exact) <= 1e-5 * max(scale, 1.0)
            assert abs(b - exact) <= 1e-5 * max(scale, 1.0)


def test_distance_known_answers():
    """reduce_add known answers of test_distances.cpp:84-100 (1..8 -> 36, 1..4 -> 10), as inner products"""
    ones8, v8 = np.ones(8, np.float32), np.arange(1, 9, dtype=np.float32)
    assert port.distance(ones8, v8, port.IP) == 1.0 - 36.0
    assert port.distance(np.ones(4, np.float32), np.arange(1, 5, dtype=np.float32), port.IP) == 1.0 - 10.0
    assert port.distance(np.zeros(8, np.float32), v8, port.L2) == 204.0


@pytest.mark.parametrize("dt", [np.uint8, np.int8])
def test_integer_distances_exact(dt):
    """defaultSquaredL2 / defaultInnerProduct on int types (L2DistanceDispatcher.h:9-17, IPDistanceDispatcher.h:9-16);
    AVX-512 uint8 kernel vs scalar on random pairs is test_distances.cpp:47-70."""
    rng = np.random.default_rng(3)
    info = np.iinfo(dt)
    for dim in (128, 100, 64, 5):
        x = rng.integers(info.min, info.max + 1, dim).astype(dt)
        y = rng.integers(info.min, info.max + 1, dim).astype(dt)
        xi, yi = x.astype(np.int64), y.astype(np.int64)
        assert port.distance(x, y, port.L2) == float(np.sum((xi - yi) ** 2))
        assert port.distance(x, y, port.IP) == np.float32(1.0) - np.float32(np.sum(xi * yi))
    lo, hi = np.full(128, info.min, dt), np.full(128, info.max, dt)
    assert port.distance(lo, hi, port.L2) == float(128 * 255 * 255)


def test_num_initializations_must_be_positive():
    ix = port.OracleIndex(golden_index_path("l2_f32_d24"), port.L2)
    q = golden_arrays("l2_f32_d24")["queries"]
    with pytest.raises(ValueError):  # std::invalid_argument, Index.h:847-849
        ix.search(q, 10, 50, num_initializations=0)


def test_bruteforce_oracle_vs_float64_scan():
    case = "l2_f32_d24"
    ix = port.OracleIndex(golden_index_path(case), port.L2)
    q = golden_arrays(case)["queries"][:16]
    d, l = ix.bruteforce(q, 10)
    x = ix.vectors().astype(np.float64)
    for i in range(q.shape[0]):
        exact = np.sum((x - q[i].astype(np.float64)) ** 2, axis=1)
        order = np.lexsort((np.arange(exact.size), exact))[:10]
        assert set(order.tolist()) == set(ix.labels()[l[i]].tolist()) or np.allclose(np.sort(exact[order]), d[i], rtol=1e-5)
        assert rel_err(d[i], exact[order]) <= 1e-5
