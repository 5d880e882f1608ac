#!/usr/bin/env python
"""Concurrent single-query callers: `Index::search` fanned over T host threads through the C++ shim
(tests/cpp/concurrent_search.cpp: the reference's executeInParallel pattern, util/Multithreading.h:18-48) on the cfg1
graph, beside the unmodified reference doing the same fan-out with the same number of threads (oracle/_ref).

    python tools/concurrency_probe.py [--threads 16] [--q 8192] [--out profiles/r2_concurrency_cfg1.json]
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", default="1,4,16,64")
    ap.add_argument("--q", type=int, default=8192)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--ef", type=int, default=100)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from flatnav_b200 import synthetic
    from oracle import refbin
    from tools.workload import ensure_index
    path, binfo = ensure_index("latent", 1_000_000, 128, "l2", 32, 100, builder="gpu")
    q = synthetic.make("latent", args.q, 128, queries=True)
    rows = []
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "concurrent_search")
        subprocess.run(["g++", "-std=c++17", "-O2", "-pthread", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "cpp", "concurrent_search.cpp"), "-o", exe,
                        "-L" + os.path.join(ROOT, "flatnav_b200"), "-lflatnav_b200",
                        "-Wl,-rpath," + os.path.join(ROOT, "flatnav_b200")], check=True)
        qp = os.path.join(td, "q.bin")
        q.tofile(qp)
        for t in [int(x) for x in args.threads.split(",")]:
            r = subprocess.run([exe, path, qp, str(args.q), str(args.k), str(args.ef), str(t)], capture_output=True, text=True)
            m = re.search(r"1 thread: (\d+) queries/s, (\d+) threads: (\d+) queries/s", r.stdout)
            row = {"threads": t, "ok": r.returncode == 0 and "concurrent == serial" in r.stdout,
                   "qps_1_thread": int(m.group(1)) if m else None, "qps_threads": int(m.group(3)) if m else None}
            if refbin.available() and t <= (os.cpu_count() or 1):
                _, _, info = refbin.search(path, "l2", q[:4000], args.k, args.ef, 100, threads=t, reps=2, want_results=False)
                row["reference_qps_same_threads"] = info["qps_best"]
            rows.append(row)
            print(json.dumps(row), flush=True)
    out = {"workload": "cfg1 1Mx128 f32 L2, K=%d ef=%d, one Index::search call per query" % (args.k, args.ef),
           "host_cores": os.cpu_count(), "index_build": binfo, "rows": rows}
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
