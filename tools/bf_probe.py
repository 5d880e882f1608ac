"""Development probe: time fnb_bruteforce (tensor path vs CUDA-core exact scan) on a raw synthetic index.
    python tools/bf_probe.py [N] [D] [Q] [K] [gen] [metric]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flatnav_b200  # noqa: E402
from flatnav_b200 import synthetic  # noqa: E402
from tools.rawindex import index_bytes  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
Q = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
K = int(sys.argv[4]) if len(sys.argv) > 4 else 10
gen = sys.argv[5] if len(sys.argv) > 5 else "latent"
metric = sys.argv[6] if len(sys.argv) > 6 else "l2"
modes = (sys.argv[7] if len(sys.argv) > 7 else "tensor,exact").split(",")

data = synthetic.make(gen, N, D)
q = synthetic.make(gen, Q, D, queries=True)
cls = {("l2", "float32"): "IndexL2Float", ("ip", "float32"): "IndexIPFloat", ("l2", "uint8"): "IndexL2Uint8",
       ("ip", "uint8"): "IndexIPUint8", ("l2", "int8"): "IndexL2Int8", ("ip", "int8"): "IndexIPInt8"}[(metric, data.dtype.name)]
ix = getattr(flatnav_b200.index, cls).from_bytes(index_bytes(data, M=4))
res = {}
for mode in modes:
    os.environ["FNB_BF_MODE"] = mode
    for rep in range(2):
        t0 = time.time()
        d, l = ix.bruteforce(q, K)
        wall = time.time() - t0
    st = ix.last_bruteforce_stats
    res[mode] = (d, l)
    tf = st["gemm_flops"] / (st["gemm_ms"] * 1e-3) / 1e12 if st["gemm_ms"] else 0.0
    print(f"{mode}: N={N} D={D} Q={Q} K={K} {gen}/{metric} wall={wall*1e3:.1f} ms stats={st} gemm_tflops={tf:.1f}", flush=True)
if len(res) == 2:
    a, b = res["tensor"], res["exact"]
    print("bit-identical:", np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1]))
