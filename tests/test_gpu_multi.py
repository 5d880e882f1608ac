"""GPU (-m gpu), needs >= 2 devices (skipped otherwise): replicas inside one process, and one process per GPU over
NCCL for dataset sharding (all-gather of per-shard top-K + fnb_merge_topk on the device)."""
import os
import socket
import sys

import numpy as np
import pytest

import flatnav_b200
from conftest import ROOT, golden_arrays, golden_index_path
from oracle import port

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_replicated_query_sharding_in_one_process():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    case = "l2_f32_d24"
    q = golden_arrays(case)["queries"]
    one = flatnav_b200.index.IndexL2Float.load_index(golden_index_path(case), devices=[0])
    two = flatnav_b200.index.IndexL2Float.load_index(golden_index_path(case), devices=[0, 1])
    assert two.info["n_devices"] == 2 and two.info["device_ids"] == [0, 1]
    for K, ef in ((10, 50), (1, 16)):
        d1, l1 = one.search(q, K, ef)
        d2, l2 = two.search(q, K, ef)
        np.testing.assert_array_equal(d1, d2)
        np.testing.assert_array_equal(l1, l2)
        assert two.last_stats["n_dist"] == one.last_stats["n_dist"]
    d2, l2 = two.search(q[:1], 10, 50)  # fewer queries than replicas
    np.testing.assert_array_equal(d2, one.search(q[:1], 10, 50)[0])


def _shard_file(path, vectors, links, labels):
    n, d = vectors.shape
    M = links.shape[1]
    blob = bytearray()
    blob += np.int32(9).tobytes() + np.array([M, 4 * d, 4 * d + 4 * M + 4, n, n, d, 4 * d], dtype=np.uint64).tobytes()
    for i in range(n):
        blob += vectors[i].tobytes() + links[i].tobytes() + np.int32(labels[i]).tobytes()
    open(path, "wb").write(bytes(blob))


def _worker(rank, world, port_no, out_dir, shard_paths):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from flatnav_b200.distributed import DatasetShardedSearcher
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    q = np.load(os.path.join(ROOT, "tests", "golden", "l2_f32_d24.npz"))["queries"]
    ix = flatnav_b200.index.IndexL2Float.load_index(shard_paths[rank], devices=[rank])
    for mode in ("nccl", "peer"):
        sh = DatasetShardedSearcher(ix, exchange=mode, max_queries=4096, max_k=16)
        for rep in range(3):  # several epochs through the double-buffered peer exchange
            d, l = sh.search(q, 10, 50)
        d1, l1 = sh.search(q[:7], 3, 20)  # a different (Q, K) through the same buffers
        np.save(os.path.join(out_dir, f"d_{mode}_{rank}.npy"), d)
        np.save(os.path.join(out_dir, f"l_{mode}_{rank}.npy"), l)
        np.save(os.path.join(out_dir, f"d1_{mode}_{rank}.npy"), d1)
        np.save(os.path.join(out_dir, f"l1_{mode}_{rank}.npy"), l1)
        sh.close()
    dist.destroy_process_group()


def test_dataset_sharding_over_nccl(tmp_path):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    full = port.OracleIndex(golden_index_path("l2_f32_d24"), port.L2)
    vec = full.vectors()
    n = vec.shape[0]
    paths = []
    for r, (a, b) in enumerate([(0, n // 2), (n // 2, n)]):
        m = b - a
        idx = np.arange(m)
        links = np.stack([(idx + k) % m for k in (1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233)], axis=1).astype(np.uint32)
        p = str(tmp_path / f"shard{r}.idx")
        _shard_file(p, vec[a:b], links, np.arange(a, b))
        paths.append(p)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_no = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port_no, str(tmp_path), paths), nprocs=2, join=True)
    q = golden_arrays("l2_f32_d24")["queries"]
    per = [port.OracleIndex(p, port.L2).search(q, 10, 50, mode=port.MODE_LIST) for p in paths]
    exp_d = np.empty((q.shape[0], 10), np.float32)
    exp_l = np.empty((q.shape[0], 10), np.int32)
    for i in range(q.shape[0]):
        pairs = sorted((float(per[s][0][i, k]), int(per[s][1][i, k])) for s in range(2) for k in range(10))[:10]
        exp_d[i] = [p[0] for p in pairs]
        exp_l[i] = [p[1] for p in pairs]
    for mode in ("nccl", "peer"):
        for r in range(2):
            np.testing.assert_array_equal(np.load(tmp_path / f"d_{mode}_{r}.npy"), exp_d)
            np.testing.assert_array_equal(np.load(tmp_path / f"l_{mode}_{r}.npy"), exp_l)
            np.testing.assert_array_equal(np.load(tmp_path / f"d1_{mode}_{r}.npy"), np.load(tmp_path / "d1_nccl_0.npy"))
            np.testing.assert_array_equal(np.load(tmp_path / f"l1_{mode}_{r}.npy"), np.load(tmp_path / "l1_nccl_0.npy"))
