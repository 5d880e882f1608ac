#!/usr/bin/env python
"""Dataset-sharded search benchmark (BASELINE.json cfg5 shape: uint8 D=128 L2, M=32, one sub-graph per GPU, labels
= global ids, every rank answers all queries, global top-K by peer-memory exchange + on-device merge).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_sharded.py [--n-shard 12500000] [--q 10000] [--ef 100] [--steps 50]

Each rank builds its own shard: --builder gpu (default) uses this engine's GPU construction (csrc/build.cu, a few
seconds for 12.5M x 128 uint8, so --n-shard 12500000 gives the literal 100M / 8 configuration), --builder ref the
unmodified reference on this rank's share of the host cores.  Prints one JSON line per exchange mode (rank 0): QPS (queries/s of the whole
job: all ranks answer the same Q), ms per step (CUDA events, max over ranks), recall@K against the exact global
ground truth (per-shard tensor-core brute force, merged)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-shard", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--gen", default="latent-u8")
    ap.add_argument("--q", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--ef", type=int, default=100)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--builder", default="gpu", choices=["gpu", "ref"])
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import flatnav_b200
    from flatnav_b200 import synthetic
    from flatnav_b200.distributed import DatasetShardedSearcher, merge_topk_cuda
    from oracle import refbin
    from tools.workload import CACHE

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    cls = {"latent-u8": flatnav_b200.index.IndexL2Uint8, "latent": flatnav_b200.index.IndexL2Float,
           "latent-i8": flatnav_b200.index.IndexL2Int8}[args.gen]
    t0 = time.time()
    data = synthetic.make(args.gen, args.n_shard, args.dim, stream=rank + 1)
    gen_s = time.time() - t0
    t0 = time.time()
    if args.builder == "gpu":
        from flatnav_b200.data_type import DataType
        dt = {"uint8": DataType.uint8, "float32": DataType.float32, "int8": DataType.int8}[data.dtype.name]
        ix = flatnav_b200.index.create("l2", args.dim, args.n_shard, 32, dt)
        ix.add(data, 100, labels=np.arange(rank * args.n_shard, (rank + 1) * args.n_shard, dtype=np.int32))
        build_info = dict(ix.last_build_stats)
    else:
        os.makedirs(CACHE, exist_ok=True)
        path = os.path.join(CACHE, f"shard{rank}of{world}_{args.gen}_n{args.n_shard}_d{args.dim}_l2_M32_efc100.idx")
        threads = max(1, (os.cpu_count() or 1) // world)
        build_info = refbin.build_index(data, "l2", 32, 100, path, threads=threads, first_label=rank * args.n_shard)
        ix = cls.load_index(path, devices=[local])
    del data
    build_s = time.time() - t0
    q = synthetic.make(args.gen, args.q, args.dim, queries=True)
    Q, K = args.q, args.k
    dq = torch.from_numpy(q).cuda()

    # exact global ground truth: per-shard exact scan (tensor-core path), gathered and merged
    gd_, gl_ = ix.bruteforce(q, K)
    td, tl = torch.from_numpy(gd_).cuda(), torch.from_numpy(gl_).cuda()
    ad = torch.empty((world * Q, K), dtype=torch.float32, device="cuda")
    al = torch.empty((world * Q, K), dtype=torch.int32, device="cuda")
    dist.all_gather_into_tensor(ad, td)
    dist.all_gather_into_tensor(al, tl)
    _, gt = merge_topk_cuda(ad.view(world, Q, K), al.view(world, Q, K), K)
    gt = gt.cpu().numpy()

    results = {}
    for mode in ("nccl", "peer"):
        sh = DatasetShardedSearcher(ix, exchange=mode, max_queries=Q, max_k=K)
        for _ in range(args.warmup):
            od, ol = sh.search_tensors(dq, K, args.ef)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            od, ol = sh.search_tensors(dq, K, args.ef)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / args.steps
        lab = ol.cpu().numpy()
        rec = float(np.mean([len(set(a.tolist()) & set(b.tolist())) / K for a, b in zip(lab, gt)]))
        results[mode] = (od.cpu().numpy(), lab)
        if rank == 0:
            print(json.dumps({"bench": "dataset-sharded search", "exchange": mode, "n_gpus": world,
                              "n_total": world * args.n_shard, "n_shard": args.n_shard, "dim": args.dim, "gen": args.gen,
                              "Q": Q, "K": K, "ef_search": args.ef, "steps": args.steps, "ms_per_step": ms,
                              "qps": Q / (ms * 1e-3), "recall_at_k": round(rec, 4), "builder": args.builder, "shard_gen_s": round(gen_s, 1),
                              "shard_build_s": round(build_s, 1), "shard_build": build_info}),
                  flush=True)
        sh.close()
    same = np.array_equal(results["nccl"][0], results["peer"][0]) and np.array_equal(results["nccl"][1], results["peer"][1])
    if rank == 0:
        print(json.dumps({"peer_equals_nccl": bool(same)}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
