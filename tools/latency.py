#!/usr/bin/env python
"""Per-query latency of `search_single`, the protocol of the reference's benchmark driver.

Restates `compute_metrics` of experiments/run-benchmark.py:38-124 and the metric definitions of
experiments/plotting/metrics.py:53-132 for this engine: every query goes through `index.search_single(query, K,
ef_search, num_initializations=100)` with `time.time()` around the call, then

    recall                 mean over queries of |returned ∩ ground truth| / K
    qps                    num_queries / sum(latencies)
    latency_p50/p90/p95/p99/p999   percentiles of the per-call wall time, in ms
    distance_computations  index.get_query_distance_computations() summed / num_queries

Beside it: the unmodified reference (oracle/_ref) doing the same loop on one host thread, timed inside its own
process (no Python in its numbers), and this engine's batched `search` over the same queries for contrast.

    python tools/latency.py [cfg1] [--q 2000] [--efs 32,64,100,200] [--out profiles/r1_latency_cfg1.json]
    python tools/latency.py cfg1 --paper       # k = 100 and the authors' ef_search list
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

from tools.sweep import CLS, CONFIGS  # noqa: E402

PCTS = {"latency_p50": 50, "latency_p90": 90, "latency_p95": 95, "latency_p99": 99, "latency_p999": 99.9}


def compute_recall(ground_truth: np.ndarray, top_k: np.ndarray, k: int) -> float:
    sets = [set(g.tolist()) for g in ground_truth]
    return float(np.mean([sum(1 for x in row.tolist() if x in sets[i]) / k for i, row in enumerate(top_k)]))


def summarize(lat_s: np.ndarray) -> dict:
    out = {name: float(np.percentile(lat_s, p) * 1000) for name, p in PCTS.items()}
    out["latency_mean"] = float(np.mean(lat_s) * 1000)
    out["qps"] = float(len(lat_s) / np.sum(lat_s))
    return out


def compute_metrics(index, queries: np.ndarray, ground_truth: np.ndarray, ef_search: int, k: int) -> dict:
    latencies, top_k = [], []
    index.get_query_distance_computations()  # reset
    ndist = 0
    for query in queries:
        start = time.time()
        _, indices = index.search_single(query=query, ef_search=ef_search, K=k, num_initializations=100)
        end = time.time()
        latencies.append(end - start)
        top_k.append(indices)
        ndist += index.get_query_distance_computations()
    m = summarize(np.asarray(latencies))
    m["recall"] = compute_recall(ground_truth, np.asarray(top_k), k)
    m["distance_computations"] = ndist / len(queries)
    return m


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default="cfg1")
    ap.add_argument("--q", type=int, default=2000)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--efs", default="32,64,100,200")
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--builder", default="reference")
    ap.add_argument("--k", type=int, default=0, help="override the config's K")
    ap.add_argument("--paper", action="store_true",
                    help="the authors' settings (experiments/Makefile:8-22, run-benchmark.py:44): k = 100, "
                         "ef_search in {100, 200, 300, 500, 1000, 3000}")
    ap.add_argument("--time-kernels", action="store_true",
                    help="also record the device time of every search_single launch (FNB_TIME_KERNELS=1: four event "
                         "records per call, which the protocol numbers then include)")
    args = ap.parse_args()
    if args.time_kernels:
        os.environ["FNB_TIME_KERNELS"] = "1"
    if args.paper:
        args.k = args.k or 100
        args.efs = "100,200,300,500,1000,3000"

    import flatnav_b200
    from flatnav_b200 import synthetic
    from oracle import refbin
    from tools.workload import ensure_index

    c = dict(CONFIGS[args.config])
    if args.n:
        c["n"] = args.n
    rank = c.get("rank", 16)
    path, binfo = ensure_index(c["gen"], c["n"], c["dim"], c["metric"], c["M"], c["efc"], rank=rank, builder=args.builder)
    queries = synthetic.make(c["gen"], args.q, c["dim"], queries=True, rank=rank)
    cls = getattr(flatnav_b200.index, CLS[(c["metric"], queries.dtype.name)])
    ix = cls.load_index(path, devices=[0])
    K = args.k or c["K"]
    _, gt = ix.bruteforce(queries, K)
    for q in queries[:200]:  # warm-up: context, workspace, clocks
        ix.search_single(q, K, 64)

    rows = []
    for ef in [int(x) for x in args.efs.split(",")]:
        ef = max(ef, 1)
        single = compute_metrics(ix, queries, gt, ef, K)
        # kernel-only view of one query: device time between the events fnb_search records around its launch
        if args.time_kernels:
            kms = []
            for q in queries[:500]:
                ix.search_single(q, K, ef)
                kms.append(ix.last_stats["kernel_ms"])
            single["kernel_ms_p50"] = float(np.percentile(kms, 50))
        t0 = time.time()
        _, lab = ix.search(queries, K, ef)
        batched = {"qps": len(queries) / (time.time() - t0), "recall": compute_recall(gt, lab, K)}
        row = {"ef": ef, "search_single": single, "batched_search": batched}
        if not args.no_ref and refbin.available():
            lat = refbin.latency(path, c["metric"], queries, K, ef)
            row["reference_1thread"] = summarize(lat)
        rows.append(row)
        print(json.dumps(row), flush=True)
    out = {"config": args.config, "params": dict({k: v for k, v in c.items() if k != "efs"}, K=K), "num_queries": args.q,
           "protocol": "experiments/run-benchmark.py compute_metrics: search_single per query, time.time() around each call",
           "host_cores": os.cpu_count(), "reference_isa": refbin.isa(), "index_build": binfo, "rows": rows}
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
