#!/usr/bin/env python
"""Profiling target (not the bench contract): a few launches of ONE traversal-kernel shape between
cudaProfilerStart / cudaProfilerStop, so that `ncu --profile-from-start off` captures the shipped kernel on a BASELINE
shape without the index construction (which launches the same kernel template) getting in the way.

    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fnb_search_kernel -c 1 \
        -f -o gpurun_out/prof_u8 python tools/ncu_one.py u8 [--ef 100] [--q 10000] [--launches 2]

Cases are tools/ab_probe.py's (cfg1, cfg2, u8, cfg3s, cfg4s, cfg1big, ...).  FNB_LIB_PATH selects a variant build.
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> None:
    from tools.ab_probe import CASES
    ap = argparse.ArgumentParser()
    ap.add_argument("case", choices=sorted(CASES))
    ap.add_argument("--ef", type=int, default=100)
    ap.add_argument("--q", type=int, default=0)
    ap.add_argument("--launches", type=int, default=2)
    ap.add_argument("--single", action="store_true", help="profile search_single calls (latency variant) instead")
    args = ap.parse_args()

    import torch

    import flatnav_b200
    from flatnav_b200 import synthetic
    from flatnav_b200.data_type import DataType
    from tools.workload import ensure_index
    c = CASES[args.case]
    Q, K = args.q or c["Q"], c["K"]
    path, _ = ensure_index(c["gen"], c["n"], c["dim"], c["metric"], 32, 100, rank=c.get("rank", 16), builder="gpu")
    queries = synthetic.make(c["gen"], Q, c["dim"], queries=True, rank=c.get("rank", 16))
    dt = {"float32": DataType.float32, "uint8": DataType.uint8, "int8": DataType.int8}[queries.dtype.name]
    ix = flatnav_b200.index.index_class("l2" if c["metric"] == "l2" else "angular", dt).load_index(path, devices=[0])
    if args.single:
        for i in range(8):
            ix.search_single(queries[i], K, args.ef)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for i in range(8, 8 + args.launches):
            ix.search_single(queries[i], K, args.ef)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    dq = torch.from_numpy(queries).cuda()
    od = torch.empty((Q, K), dtype=torch.float32, device="cuda")
    ol = torch.empty((Q, K), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        ix.search_device(dq.data_ptr(), Q, K, args.ef, 100, od.data_ptr(), ol.data_ptr(), stream)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(args.launches):
        ix.search_device(dq.data_ptr(), Q, K, args.ef, 100, od.data_ptr(), ol.data_ptr(), stream)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    nd, nh, _ = ix.device_totals()
    print(f"ncu_one {args.case} ef={args.ef} Q={Q}: n_dist/q={nd / Q:.1f} n_hops/q={nh / Q:.2f}", flush=True)


if __name__ == "__main__":
    main()
