"""CPU, world_size 2, gloo: the host-side logic of the two multi-GPU modes (flatnav_b200/distributed.py) —
query partitioning + gather order, and dataset sharding's all-gather + k-way merge order — with the oracle standing
in for the CUDA search and a numpy merge standing in for fnb_merge_topk (both checkers, tests only)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, golden_arrays, golden_index_path
from flatnav_b200.distributed import DatasetShardedSearcher, QueryShardedSearcher, partition
from oracle import port


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partition_covers_everything_in_order():
    for n in (0, 1, 7, 10000, 100003):
        for world in (1, 2, 3, 8):
            spans = [partition(n, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == n
            pos = 0
            for s, c in spans:
                assert s == min(pos, n) and c >= 0
                pos = s + c


class _OracleAsIndex:
    """quacks like flatnav_b200.index._GpuIndex.search, backed by the CPU oracle"""

    def __init__(self, path, metric):
        self.ix = port.OracleIndex(path, metric)

    def search(self, q, K, ef, ninit=100):
        return self.ix.search(q, K, ef, ninit, mode=port.MODE_LIST)


def numpy_merge(gd, gl, K):
    gd, gl = gd.numpy(), gl.numpy()
    S, Q, _ = gd.shape
    od = np.full((Q, K), np.inf, np.float32)
    ol = np.full((Q, K), -1, np.int32)
    for q in range(Q):
        pairs = sorted((float(gd[s, q, k]), int(gl[s, q, k])) for s in range(S) for k in range(K) if gl[s, q, k] >= 0)
        for j, (d, l) in enumerate(pairs[:K]):
            od[q, j], ol[q, j] = d, l
    return torch.from_numpy(od), torch.from_numpy(ol)


def _worker(rank, world, port_no, out_dir, shard_paths):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(ROOT, "tests", "golden", "l2_f32_d24.npz"))
    q = g["queries"][:37]  # odd count: uneven slices
    # --- query sharding over a replicated index
    rep = _OracleAsIndex(os.path.join(ROOT, "tests", "golden", "l2_f32_d24.idx"), port.L2)
    d, l = QueryShardedSearcher(rep).search(q, 10, 50)
    np.save(os.path.join(out_dir, f"qs_d_{rank}.npy"), d)
    np.save(os.path.join(out_dir, f"qs_l_{rank}.npy"), l)
    # --- dataset sharding: this rank's sub-graph, all queries
    shard = port.OracleIndex(shard_paths[rank], port.L2)
    s = DatasetShardedSearcher(None, search_fn=lambda qq, K, ef, ni: shard.search(qq, K, ef, ni, mode=port.MODE_LIST),
                               merge_fn=numpy_merge)
    d2, l2 = s.search(q, 10, 50)
    np.save(os.path.join(out_dir, f"ds_d_{rank}.npy"), d2)
    np.save(os.path.join(out_dir, f"ds_l_{rank}.npy"), l2)
    dist.destroy_process_group()


def _make_shard_file(path, vectors, links, labels):
    n, d = vectors.shape
    M = links.shape[1]
    blob = bytearray()
    blob += np.int32(9).tobytes() + np.array([M, 4 * d, 4 * d + 4 * M + 4, n, n, d, 4 * d], dtype=np.uint64).tobytes()
    for i in range(n):
        blob += vectors[i].tobytes() + links[i].tobytes() + np.int32(labels[i]).tobytes()
    open(path, "wb").write(bytes(blob))


def test_world_size_2_gloo(tmp_path):
    # two shards over contiguous id ranges of the golden dataset, labels = global ids; links: a ring + skips so
    # every node is reachable (construction is out of scope; the graph only has to be a valid index file)
    full = port.OracleIndex(golden_index_path("l2_f32_d24"), port.L2)
    vec = full.vectors()
    n = vec.shape[0]
    halves = [(0, n // 2), (n // 2, n)]
    paths = []
    for r, (a, b) in enumerate(halves):
        m = b - a
        idx = np.arange(m)
        links = np.stack([(idx + k) % m for k in (1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233)], axis=1).astype(np.uint32)
        p = str(tmp_path / f"shard{r}.idx")
        _make_shard_file(p, vec[a:b], links, np.arange(a, b))
        paths.append(p)
    port_no = _free_port()
    mp.spawn(_worker, args=(2, port_no, str(tmp_path), paths), nprocs=2, join=True)

    q = golden_arrays("l2_f32_d24")["queries"][:37]
    # query sharding == single-process result, identical on both ranks
    d_ref, l_ref = full.search(q, 10, 50, mode=port.MODE_LIST)
    for r in range(2):
        np.testing.assert_array_equal(np.load(tmp_path / f"qs_d_{r}.npy"), d_ref)
        np.testing.assert_array_equal(np.load(tmp_path / f"qs_l_{r}.npy"), l_ref)
    # dataset sharding == merge of the per-shard results, identical on both ranks, labels are global ids
    per = [port.OracleIndex(p, port.L2).search(q, 10, 50, mode=port.MODE_LIST) for p in paths]
    gd = torch.from_numpy(np.stack([x[0] for x in per]))
    gl = torch.from_numpy(np.stack([x[1] for x in per]))
    md, ml = numpy_merge(gd, gl, 10)
    for r in range(2):
        np.testing.assert_array_equal(np.load(tmp_path / f"ds_d_{r}.npy"), md.numpy())
        np.testing.assert_array_equal(np.load(tmp_path / f"ds_l_{r}.npy"), ml.numpy())
    assert ml.numpy().max() >= n // 2  # results really come from both shards
    assert np.all(np.diff(md.numpy(), axis=1) >= 0)
