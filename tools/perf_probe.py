#!/usr/bin/env python
"""Development probe (not the bench contract): QPS / roofline / recall sweep on one GPU.

    python tools/perf_probe.py --n 1000000 --dim 128 --gen latent --metric l2 --ef 32,64,100 [--cpu]
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
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--gen", default="latent")
    ap.add_argument("--metric", default="l2")
    ap.add_argument("--M", type=int, default=32)
    ap.add_argument("--efc", type=int, default=100)
    ap.add_argument("--q", type=int, default=10000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--ef", default="16,32,64,100,200")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu", action="store_true", help="also time the reference CPU search (oracle/_ref)")
    ap.add_argument("--gt", action="store_true", help="compute recall against GPU brute force")
    args = ap.parse_args()

    import torch
    import flatnav_b200
    from flatnav_b200.data_type import DataType

    path, build_info = ensure_index(args.gen, args.n, args.dim, args.metric, args.M, args.efc)
    print(json.dumps({"index": path, **build_info}), flush=True)
    queries = synthetic.make(args.gen, args.q, args.dim, queries=True)
    dt = {np.dtype(np.float32): DataType.float32, np.dtype(np.uint8): DataType.uint8, np.dtype(np.int8): DataType.int8}[queries.dtype]
    cls = flatnav_b200.index.index_class("l2" if args.metric == "l2" else "angular", dt)
    t0 = time.time()
    ix = cls.load_index(path, devices=[0])
    print(json.dumps({"load_s": round(time.time() - t0, 2), "info": ix.info}), flush=True)

    gt = None
    if args.gt:
        t0 = time.time()
        _, gt = ix.bruteforce(queries, args.k)
        print(json.dumps({"bruteforce_s": round(time.time() - t0, 2)}), flush=True)

    dq = torch.from_numpy(queries).cuda()
    od = torch.empty((args.q, args.k), dtype=torch.float32, device="cuda")
    ol = torch.empty((args.q, args.k), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    peak = 6532.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for ef in [int(x) for x in args.ef.split(",")]:
        for _ in range(2):
            ix.search_device(dq.data_ptr(), args.q, args.k, ef, 100, od.data_ptr(), ol.data_ptr(), stream)
        torch.cuda.synchronize()
        times = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ix.search_device(dq.data_ptr(), args.q, args.k, ef, 100, od.data_ptr(), ol.data_ptr(), stream)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        nd, nh, ns = ix.device_totals()
        ms = float(np.median(times))
        info = ix.info
        bytes_ = nd * info["data_size_bytes"] + nh * info["max_edges_per_node"] * 4 + args.q * info["data_size_bytes"] + args.q * args.k * 8
        row = {"ef": ef, "ms": round(ms, 3), "qps": round(args.q / ms * 1e3), "ndist_q": round(nd / args.q, 1),
               "nhops_q": round(nh / args.q, 1), "GBps": round(bytes_ / ms / 1e6, 1), "roofline_frac": round(bytes_ / ms / 1e6 / peak, 3),
               "short": ns}
        if gt is not None:
            lab = ol.cpu().numpy()
            row["recall"] = round(float(np.mean([len(set(a.tolist()) & set(b.tolist())) / args.k for a, b in zip(lab, gt)])), 4)
        # end-to-end through the host-buffer entry point
        t = []
        for _ in range(3):
            t0 = time.perf_counter()
            ix.search(queries, args.k, ef)
            t.append(time.perf_counter() - t0)
        row["e2e_qps"] = round(args.q / min(t))
        if args.cpu:
            from oracle import refbin
            nq = min(args.q, 2000)
            _, _, ci = refbin.search(path, args.metric, queries[:nq], args.k, ef, threads=os.cpu_count() or 1, reps=2, want_results=False)
            row["cpu_qps_allcores"] = round(ci["qps_best"])
            row["cpu_threads"] = ci["threads"]
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
