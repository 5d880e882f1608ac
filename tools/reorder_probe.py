#!/usr/bin/env python
"""Development probe: what graph re-ordering (gorder / rcm, csrc/reorder.cu) does to the traversal on one GPU.

    python tools/reorder_probe.py --n 1000000 --dim 128 --ef 32,100 [--ref]

For the original order and after each strategy: kernel-only QPS, n_dist / n_hops per query, recall@K, and the time
the re-ordering took (host ordering + device relabel).  --ref also times the reference's own reorder on the host.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from flatnav_b200 import synthetic  # noqa: E402
from tools.workload import ensure_index  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--gen", default="latent")
    ap.add_argument("--metric", default="l2")
    ap.add_argument("--q", type=int, default=10000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--ef", default="32,100")
    ap.add_argument("--builder", default="reference")
    ap.add_argument("--strategies", default="gorder;rcm;rcm,gorder")
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="wrap one extra launch per (order, ef) in cudaProfilerStart/Stop: under `ncu --profile-from-start "
                         "off --metrics lts__t_sector_hit_rate.pct,dram__bytes_read.sum,gpu__time_duration.sum` the "
                         "captured launches are those rows, in order (tools/reorder_merge_ncu.py joins them)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch

    import flatnav_b200
    from flatnav_b200.data_type import DataType

    path, info = ensure_index(args.gen, args.n, args.dim, args.metric, builder=args.builder)
    queries = synthetic.make(args.gen, args.q, args.dim, queries=True)
    dt = {np.dtype(np.float32): DataType.float32, np.dtype(np.uint8): DataType.uint8, np.dtype(np.int8): DataType.int8}[queries.dtype]
    cls = flatnav_b200.index.index_class("l2" if args.metric == "l2" else "angular", dt)
    dq = torch.from_numpy(queries).cuda()
    od = torch.empty((args.q, args.k), dtype=torch.float32, device="cuda")
    ol = torch.empty((args.q, args.k), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    efs = [int(x) for x in args.ef.split(",")]
    rows = []

    def measure(ix, tag, extra):
        for ef in efs:
            for _ in range(3):
                ix.search_device(dq.data_ptr(), args.q, args.k, ef, 100, od.data_ptr(), ol.data_ptr(), stream)
            torch.cuda.synchronize()
            ts = []
            for _ in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ix.search_device(dq.data_ptr(), args.q, args.k, ef, 100, od.data_ptr(), ol.data_ptr(), stream)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            nd, nh, _ = ix.device_totals()
            lab = ol.cpu().numpy()
            rec = float(np.mean([len(set(a.tolist()) & set(b.tolist())) / args.k for a, b in zip(lab, gt)]))
            ms = float(np.median(ts))
            row = {"order": tag, "ef": ef, "kernel_ms": round(ms, 4), "qps": round(args.q / ms * 1e3),
                   "n_dist": round(nd / args.q, 1), "n_hops": round(nh / args.q, 1), "recall": round(rec, 4), **extra}
            rows.append(row)
            print(json.dumps(row), flush=True)
            if args.profile:
                torch.cuda.synchronize()
                torch.cuda.profiler.start()
                ix.search_device(dq.data_ptr(), args.q, args.k, ef, 100, od.data_ptr(), ol.data_ptr(), stream)
                torch.cuda.synchronize()
                torch.cuda.profiler.stop()

    ix = cls.load_index(path, devices=[0])
    _, gt = ix.bruteforce(queries, args.k)
    measure(ix, "original", {})
    for seq in args.strategies.split(";"):
        ix = cls.load_index(path, devices=[0])
        t0 = time.time()
        ix.reorder(seq.split(","))
        extra = {"reorder_s": round(time.time() - t0, 2)}
        if args.ref:
            from oracle import refbin
            r = refbin.reorder(path, args.metric, synthetic.dtype_code(queries), seq.split(","), path + ".reordered.tmp")
            os.remove(path + ".reordered.tmp")
            extra["reference_reorder_s"] = r["seconds"]
        measure(ix, seq, extra)
    if args.out:
        json.dump({"args": vars(args), "index_build": info, "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
