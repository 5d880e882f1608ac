"""Development probe: GPU construction (fnb_index_add) vs the reference's addBatch on the same data: build time,
degree, recall@10 per ef.   python tools/build_probe.py N D gen metric M [efs]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flatnav_b200  # noqa: E402
from flatnav_b200 import synthetic  # noqa: E402
from flatnav_b200.data_type import DataType  # noqa: E402
from oracle import refbin  # noqa: E402

N, D = int(sys.argv[1]), int(sys.argv[2])
gen, metric, M = sys.argv[3], sys.argv[4], int(sys.argv[5])
efs = [int(x) for x in (sys.argv[6] if len(sys.argv) > 6 else "32,64,100,200").split(",")]
with_ref = os.environ.get("NO_REF") is None
data = synthetic.make(gen, N, D)
q = synthetic.make(gen, 2000, D, queries=True)
dt = {np.dtype(np.float32): DataType.float32, np.dtype(np.uint8): DataType.uint8, np.dtype(np.int8): DataType.int8}[data.dtype]
t0 = time.time()
ix = flatnav_b200.index.create(metric, D, N, M, dt)
ix.add(data, 100)
t_gpu = time.time() - t0
print(f"gpu build: wall {t_gpu:.2f}s stats {ix.last_build_stats} -> {N / (ix.last_build_stats['device_ms'] * 1e-3):.0f} inserts/s", flush=True)
gt = ix.bruteforce(q, 10)[1]
rec = lambda l: float(np.mean([len(set(a.tolist()) & set(b.tolist())) / 10 for a, b in zip(l, gt)]))
if with_ref:
    m = "l2" if metric == "l2" else "ip"
    path = f"/tmp/bp_{gen}_{N}_{D}_{m}_{M}.idx"
    t0 = time.time()
    info = refbin.build_index(data, m, M, 100, path)
    print(f"ref build: {info['seconds']:.2f}s on {info['threads']} threads", flush=True)
    rx = type(ix).load_index(path)
for ef in efs:
    _, l = ix.search(q, 10, ef)
    line = f"ef={ef}: gpu-built recall {rec(l):.4f} n_dist/q {ix.last_stats['n_dist'] / 2000:.0f}"
    if with_ref:
        _, lr = rx.search(q, 10, ef)
        line += f" | ref-built recall {rec(lr):.4f} n_dist/q {rx.last_stats['n_dist'] / 2000:.0f}"
    print(line, flush=True)
