"""GPU (-m gpu): fnb_rerank / index.rerank / search(exact_rerank=True) — SURVEY.md §8f rank 4.  The reference has no
re-rank of its own; semantics are pinned against exact distances recomputed on the CPU from the stored vectors and
against the engine's own exact scan."""
import numpy as np
import pytest

import flatnav_b200
from conftest import build_ref_index, golden_arrays, golden_cases, golden_index_path
from flatnav_b200 import synthetic
from flatnav_b200.data_type import DataType
from oracle import port

pytestmark = pytest.mark.gpu

CASES = golden_cases()
PM = {"l2": port.L2, "ip": port.IP}
DT = {"f32": DataType.float32, "u8": DataType.uint8, "i8": DataType.int8}


def gpu_class(case):
    return flatnav_b200.index.index_class("l2" if case["metric"] == "l2" else "angular", DT[case["dtype"]])


def exact_f64(ora, q, nodes):
    rows = ora.vectors()[nodes].astype(np.float64)
    qq = q.astype(np.float64)
    return 1.0 - rows @ qq if ora.metric == port.IP else ((rows - qq) ** 2).sum(axis=1)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_rerank_of_all_nodes_is_the_exact_scan(case):
    """candidates = every node (as labels, shuffled, with repeats and junk): the result is fnb_bruteforce's, bit for bit"""
    g = golden_arrays(case["name"])
    path = golden_index_path(case["name"])
    ix = gpu_class(case).load_index(path)
    ora = port.OracleIndex(path, PM[case["metric"]])
    q = g["queries"][:24]
    labels = ora.labels()
    rng = np.random.default_rng(5)
    cand = np.stack([np.concatenate([rng.permutation(labels), labels[:7], [-1, -5, 2**30]]) for _ in range(q.shape[0])])
    K = 10
    d, l = ix.rerank(q, cand, K)
    db, lb = ix.bruteforce(q, K)
    np.testing.assert_array_equal(d.view(np.uint32), db.view(np.uint32))
    np.testing.assert_array_equal(l, lb)
    # node ids instead of labels
    nodes = np.argsort(labels)  # label -> node for these indexes (labels are a permutation of the node ids)
    d2, l2 = ix.rerank(q, np.stack([rng.permutation(len(labels)) for _ in range(q.shape[0])]), K, candidates_are_labels=False)
    np.testing.assert_array_equal(d2.view(np.uint32), db.view(np.uint32))
    np.testing.assert_array_equal(l2, lb)
    # against float64 / integer arithmetic on the CPU
    for i in range(q.shape[0]):
        ex = exact_f64(ora, q[i], nodes[l[i]])
        if case["dtype"] == "f32":
            assert np.max(np.abs(d[i] - ex) / np.maximum(np.abs(ex), 1e-6)) <= 1e-5
        else:
            np.testing.assert_array_equal(d[i].astype(np.float64), ex)
        assert np.all(np.diff(d[i]) >= 0)


def test_rerank_subset_short_lists_and_errors():
    case = CASES[0]
    g = golden_arrays(case["name"])
    ix = gpu_class(case).load_index(golden_index_path(case["name"]))
    ora = port.OracleIndex(golden_index_path(case["name"]), PM[case["metric"]])
    q = g["queries"][:16]
    labels = ora.labels()
    nodes = np.argsort(labels)
    cand = np.tile(labels[:5], (q.shape[0], 1))
    d, l = ix.rerank(q, cand, 8)  # fewer candidates than K: the tail is +inf / -1
    assert np.all(np.isinf(d[:, 5:])) and np.all(l[:, 5:] == -1)
    for i in range(q.shape[0]):
        ex = exact_f64(ora, q[i], nodes[labels[:5]])
        order = np.lexsort((nodes[labels[:5]], ex))
        np.testing.assert_array_equal(l[i, :5], labels[:5][order])
    with pytest.raises(ValueError):
        ix.rerank(q[:, :-1], cand, 3)
    with pytest.raises(ValueError):
        ix.rerank(q, cand[:3], 3)


def test_rerank_follows_relabelling_and_exact_rerank_option_is_identity(ref_cache):
    path = build_ref_index(ref_cache, "l2", "latent", 20000, 128, 32, 100)
    q = synthetic.make("latent", 200, 128, queries=True)
    ix = flatnav_b200.index.IndexL2Float.load_index(path)
    d0, l0 = ix.search(q, 10, 100)
    d1, l1 = ix.search(q, 10, 100, exact_rerank=True)
    np.testing.assert_array_equal(d0.view(np.uint32), d1.view(np.uint32))
    np.testing.assert_array_equal(l0, l1)
    # after a re-ordering labels no longer equal node ids: the label table must be rebuilt, results must not move
    _, wide = ix.search(q, 100, 100)
    before = ix.rerank(q, wide, 10)
    ix.reorder(["rcm"])
    after = ix.rerank(q, wide, 10)
    np.testing.assert_array_equal(before[0].view(np.uint32), after[0].view(np.uint32))
    np.testing.assert_array_equal(before[1], after[1])
    np.testing.assert_array_equal(before[1], l0)
