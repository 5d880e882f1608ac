#!/usr/bin/env python
"""ef_search sweep of one BASELINE.json config on one GPU: for every ef record recall@K against
exact ground truth (fnb_bruteforce, tensor-core path), kernel-only QPS (CUDA events, queries resident in HBM),
n_dist / n_hops per query, algorithmic bytes per query, achieved GB/s and its fraction of the HBM peak, and the
unmodified reference's QPS on this host's cores (oracle/_ref, bounded sample).

    python tools/sweep.py cfg2 [--out profiles/r1_sweep_cfg2.json] [--n N] [--q Q] [--no-ref]

The graph is built by the unmodified reference (default) or by this engine (--builder gpu) and cached under data_cache/.
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

CONFIGS = {
    # BASELINE.json configs[0..4]
    "cfg1": dict(gen="latent", n=1_000_000, dim=128, metric="l2", M=32, efc=100, Q=10_000, K=10,
                 efs=[16, 32, 64, 100, 200, 400]),
    "cfg2": dict(gen="latent-norm", n=1_200_000, dim=100, metric="ip", M=32, efc=100, Q=10_000, K=10,
                 efs=[16, 32, 64, 128, 256, 512]),
    "cfg3": dict(gen="latent", n=10_000_000, dim=96, metric="l2", M=32, efc=100, Q=100_000, K=10,
                 efs=[32, 64, 100, 200]),
    "cfg4": dict(gen="latent", n=1_000_000, dim=960, metric="l2", M=32, efc=100, Q=10_000, K=100,
                 efs=[100, 200, 300, 400, 512], rank=32),
    "cfg5shard": dict(gen="latent-u8", n=12_500_000, dim=128, metric="l2", M=32, efc=100, Q=10_000, K=10,
                      efs=[32, 64, 100, 200]),
}
CLS = {("l2", "float32"): "IndexL2Float", ("ip", "float32"): "IndexIPFloat", ("l2", "uint8"): "IndexL2Uint8",
       ("ip", "uint8"): "IndexIPUint8", ("l2", "int8"): "IndexL2Int8", ("ip", "int8"): "IndexIPInt8"}


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--out", default=None)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--q", type=int, default=0)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--ref-sample", type=int, default=2000)
    ap.add_argument("--efs", default="", help="comma-separated ef_search list overriding the config's")
    ap.add_argument("--builder", default="reference", choices=["reference", "gpu"],
                    help="who builds the graph when it is not cached: the unmodified reference, or this engine's GPU construction")
    args = ap.parse_args()
    c = dict(CONFIGS[args.config])
    if args.n:
        c["n"] = args.n
    if args.q:
        c["Q"] = args.q
    if args.efs:
        c["efs"] = [int(x) for x in args.efs.split(",")]

    import torch

    import flatnav_b200
    from flatnav_b200 import synthetic
    from oracle import refbin
    from tools.workload import ensure_index

    kw = {"rank": c["rank"]} if "rank" in c else {}
    t0 = time.time()
    path, binfo = ensure_index(c["gen"], c["n"], c["dim"], c["metric"], c["M"], c["efc"], builder=args.builder, **kw)
    print(f"[sweep] index ready in {time.time() - t0:.1f}s: {binfo}", flush=True)
    q = synthetic.make(c["gen"], c["Q"], c["dim"], queries=True, **kw)
    cls = getattr(flatnav_b200.index, CLS[(c["metric"], q.dtype.name)])
    ix = cls.load_index(path)
    info = ix.info
    Q, K = c["Q"], c["K"]
    t0 = time.time()
    _, gt = ix.bruteforce(q, K)
    bf = dict(ix.last_bruteforce_stats)
    print(f"[sweep] ground truth in {time.time() - t0:.2f}s: {bf}", flush=True)

    dq = torch.from_numpy(q).cuda()
    dd = torch.empty((Q, K), dtype=torch.float32, device="cuda")
    dl = torch.empty((Q, K), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    pk, pk_src = peak()
    rows = []
    cores = os.cpu_count() or 1
    for ef in c["efs"]:
        for _ in range(3):
            ix.search_device(dq.data_ptr(), Q, K, ef, 100, dd.data_ptr(), dl.data_ptr(), stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            ix.search_device(dq.data_ptr(), Q, K, ef, 100, dd.data_ptr(), dl.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters  # launches back to back: adjacent launches overlap their tails (PDL)
        iso = []
        for _ in range(5):  # one launch at a time on an idle GPU
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ix.search_device(dq.data_ptr(), Q, K, ef, 100, dd.data_ptr(), dl.data_ptr(), stream)
            b.record()
            torch.cuda.synchronize()
            iso.append(a.elapsed_time(b))
        iso_ms = float(np.median(iso))
        nd, nh, ns = ix.device_totals()
        lab = dl.cpu().numpy()
        rec = float(np.mean([len(set(a.tolist()) & set(b.tolist())) / K for a, b in zip(lab, gt)]))
        algo = nd * info["data_size_bytes"] + nh * info["max_edges_per_node"] * 4 + Q * info["data_size_bytes"] + Q * K * 8
        gbs = algo / (ms * 1e-3) / 1e9
        row = dict(ef=ef, recall=round(rec, 4), qps=Q / (ms * 1e-3), kernel_ms=ms, n_dist=nd / Q, n_hops=nh / Q,
                   bytes_per_query=algo / Q, gbs=gbs, frac=gbs / pk, isolated_ms=iso_ms, isolated_qps=Q / (iso_ms * 1e-3),
                   isolated_frac=algo / (iso_ms * 1e-3) / 1e9 / pk, n_short=ns, kernel=ix.kernel_signature(Q, K, ef),
                   plan=ix.search_plan(Q, K, ef))
        if not args.no_ref and refbin.available():
            nq = min(Q, args.ref_sample)
            _, lr, rinfo = refbin.search(path, c["metric"], q[:nq], K, ef, 100, threads=cores, reps=2)
            row["ref_qps"] = rinfo["qps_best"]
            row["ref_recall"] = round(float(np.mean([len(set(a.tolist()) & set(b.tolist())) / K
                                                     for a, b in zip(lr, gt[:nq])])), 4)
            row["speedup"] = row["qps"] / row["ref_qps"]
        rows.append(row)
        print("[sweep]", json.dumps(row), flush=True)
    ok = [r for r in rows if r["recall"] >= 0.95]
    operating_point = min(ok, key=lambda r: r["ef"]) if ok else None
    out = dict(config=args.config, params={k: v for k, v in c.items() if k != "efs"}, index_build=binfo,
               operating_point_recall_095=operating_point,
               bruteforce=bf, peak_gbs=pk, peak_source=pk_src, host_cores=cores, rows=rows)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
