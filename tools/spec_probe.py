#!/usr/bin/env python
"""Development probe for the speculating latency kernel (search_cta_spec_kernel.cuh): with a library built with
-DFNB_SPEC_DEBUG (FNB_LIB_PATH), runs search_single over the same queries once per counter (FNB_SPEC_STATS=k, read by
the library at every call) and prints the per-query mean of each — events per query and clock cycles of the driver's
waits.  Not a product path: the counters replace n_dist in the output of that build.

    FNB_LIB_PATH=variants/libfnb_specdbg.so python tools/spec_probe.py [--ef 100] [--q 300]
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
NAMES = ["hops", "slot_hits", "rounds_posted", "rows_posted", "cyc_publish_candidates", "cyc_links", "cyc_visited_filter",
         "cyc_target_filter", "cyc_wait_previous_round", "cyc_slot_and_post", "cyc_wait_own_rows", "cyc_wait_list_merged",
         "cyc_next_target", "cyc_acceptance", "cyc_selection", "cyc_main_loop"]
CLASSES = {0: "all", 1: "slot_hit_hops", 2: "other_hops"}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--ef", default="100")
    ap.add_argument("--q", type=int, default=300)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import flatnav_b200
    from flatnav_b200 import synthetic
    from tools.sweep import CONFIGS
    from tools.workload import ensure_index
    c = dict(CONFIGS["cfg1"])
    path, _ = ensure_index(c["gen"], c["n"], c["dim"], c["metric"], c["M"], c["efc"], builder="gpu")
    queries = synthetic.make(c["gen"], args.q, c["dim"], queries=True)
    ix = flatnav_b200.index.IndexL2Float.load_index(path, devices=[0])
    K = c["K"]
    for q in queries[:100]:
        ix.search_single(q, K, 64)
    out = {}
    for ef in [int(x) for x in args.ef.split(",")]:
        row = {}
        for cls, cname in CLASSES.items():
            sub = {}
            for k in range(0, 17):
                if cls and k in (0, 1, 2, 3, 4, 16):
                    continue
                os.environ["FNB_SPEC_STATS"] = str(k | (cls << 8))  # low byte 0: n_dist itself
                ix.get_query_distance_computations()
                t0 = time.perf_counter()
                for q in queries:
                    ix.search_single(q, K, ef)
                dt = time.perf_counter() - t0
                sub["n_dist" if k == 0 else NAMES[k - 1]] = round(ix.get_query_distance_computations() / len(queries), 1)
                if not cls:
                    sub["ms_per_query"] = round(dt / len(queries) * 1e3, 4)
            row[cname] = sub
        out[ef] = row
        print(json.dumps({"ef": ef, **row}), flush=True)
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
