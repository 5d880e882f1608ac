#!/usr/bin/env python
"""Development A/B probe (not the bench contract): time the traversal kernel of several builds of the library on the
same graphs and queries, one subprocess per build (FNB_LIB_PATH), and check that every build returns the same bytes.

    python tools/ab_probe.py --libs base=flatnav_b200/libflatnav_b200.so,spec=variants/libspec.so \
        --cases cfg1,cfg2,u8 --out gpurun_out/ab.json
    python tools/ab_probe.py --libs auto=flatnav_b200/libflatnav_b200.so,sparse=flatnav_b200/libflatnav_b200.so@FNB_DENSE=0 ...

Graphs are built on the GPU by the default library (seconds) and cached under data_cache/.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    "cfg1": dict(gen="latent", n=1_000_000, dim=128, metric="l2", K=10, efs=[32, 64, 100, 200], Q=10_000),
    "cfg2": dict(gen="latent-norm", n=1_200_000, dim=100, metric="ip", K=10, efs=[32, 64, 128, 256], Q=10_000),
    "cfg3s": dict(gen="latent", n=4_000_000, dim=96, metric="l2", K=10, efs=[64, 100], Q=50_000),
    "cfg4s": dict(gen="latent", n=400_000, dim=960, metric="l2", K=100, efs=[100, 300], Q=5_000, rank=32),
    # large batches: the tail of the last wave of queries no longer matters (occupancy experiments)
    "cfg1big": dict(gen="latent", n=1_000_000, dim=128, metric="l2", K=10, efs=[64, 100, 200], Q=100_000),
    "cfg2big": dict(gen="latent-norm", n=1_200_000, dim=100, metric="ip", K=10, efs=[64, 128], Q=100_000),
    "u8big": dict(gen="latent-u8", n=4_000_000, dim=128, metric="l2", K=10, efs=[64, 100], Q=100_000),
    "u8": dict(gen="latent-u8", n=4_000_000, dim=128, metric="l2", K=10, efs=[32, 64, 100, 200], Q=10_000),
}


def child(case: str, path: str, reps: int) -> None:
    import torch

    import flatnav_b200
    from flatnav_b200 import synthetic
    from flatnav_b200.data_type import DataType
    c = CASES[case]
    queries = synthetic.make(c["gen"], c["Q"], c["dim"], queries=True, rank=c.get("rank", 16))
    dt = {"float32": DataType.float32, "uint8": DataType.uint8, "int8": DataType.int8}[queries.dtype.name]
    ix = flatnav_b200.index.index_class("l2" if c["metric"] == "l2" else "angular", dt).load_index(path, devices=[0])
    Q, K = c["Q"], c["K"]
    dq = torch.from_numpy(queries).cuda()
    od = torch.empty((Q, K), dtype=torch.float32, device="cuda")
    ol = torch.empty((Q, K), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    info = ix.info
    rows = []
    for ef in c["efs"]:
        for _ in range(3):
            ix.search_device(dq.data_ptr(), Q, K, ef, 100, od.data_ptr(), ol.data_ptr(), stream)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ix.search_device(dq.data_ptr(), Q, K, ef, 100, od.data_ptr(), ol.data_ptr(), stream)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        nd, nh, ns = ix.device_totals()
        ms = float(np.median(ts))
        # back-to-back launches, nothing between them on the stream (what a caller streaming batches sees)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2 * reps):
            ix.search_device(dq.data_ptr(), Q, K, ef, 100, od.data_ptr(), ol.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        stream_ms = e0.elapsed_time(e1) / (2 * reps)
        b = nd * info["data_size_bytes"] + nh * info["max_edges_per_node"] * 4 + Q * info["data_size_bytes"] + Q * K * 8
        h = hashlib.sha1(od.cpu().numpy().tobytes() + ol.cpu().numpy().tobytes()).hexdigest()[:12]
        rows.append({"ef": ef, "ms": round(ms, 4), "stream_ms": round(stream_ms, 4), "qps": round(Q / ms * 1e3), "gbs": round(b / ms / 1e6, 1),
                     "ndist_q": round(nd / Q, 1), "nhops_q": round(nh / Q, 2), "sha": h})
    print("ABROW " + json.dumps(rows), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--libs", required=True, help="name=path,... (paths relative to the repo root)")
    ap.add_argument("--cases", default="cfg1,cfg2,u8")
    ap.add_argument("--reps", type=int, default=9)
    ap.add_argument("--out", default=None)
    ap.add_argument("--child", nargs=2, default=None)
    args = ap.parse_args()
    if args.child:
        child(args.child[0], args.child[1], args.reps)
        return
    from tools.workload import ensure_index
    libs = [kv.split("=", 1) for kv in args.libs.split(",")]
    result = {}
    for case in args.cases.split(","):
        c = CASES[case]
        path, binfo = ensure_index(c["gen"], c["n"], c["dim"], c["metric"], 32, 100, rank=c.get("rank", 16), builder="gpu")
        result[case] = {"build": binfo, "libs": {}}
        for name, lp in libs:
            lp, *assign = lp.split("@")  # path@VAR=VALUE@...: the same build under different environment knobs
            env = dict(os.environ, FNB_LIB_PATH=os.path.join(ROOT, lp))
            env.update(kv.split("=", 1) for kv in assign)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--libs", "x=x", "--reps", str(args.reps),
                                "--child", case, path], env=env, capture_output=True, text=True)
            rows = None
            for line in r.stdout.splitlines():
                if line.startswith("ABROW "):
                    rows = json.loads(line[6:])
            if rows is None:
                rows = {"error": (r.stderr or r.stdout)[-400:]}
            result[case]["libs"][name] = rows
        # table
        base = result[case]["libs"][libs[0][0]]
        print(f"== {case}  (n={c['n']} dim={c['dim']} {c['metric']} K={c['K']})", flush=True)
        for name, _ in libs:
            rows = result[case]["libs"][name]
            if isinstance(rows, dict):
                print(f"  {name:10s} ERROR {rows['error']}")
                continue
            cells = []
            for r0, r1 in zip(base, rows):
                same = "=" if r0["sha"] == r1["sha"] else "DIFF"
                st = f" s{r1['stream_ms']:.3f}" if "stream_ms" in r1 else ""
                cells.append(f"ef{r1['ef']}: {r1['ms']:.3f}ms{st} x{r0['ms'] / r1['ms']:.3f} nd{r1['ndist_q']:.0f} {same}")
            print(f"  {name:10s} " + " | ".join(cells), flush=True)
    if args.out:
        json.dump(result, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
