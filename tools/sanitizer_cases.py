#!/usr/bin/env python
"""Small parity cases for `compute-sanitizer` (memcheck / racecheck / initcheck / synccheck): every traversal-kernel
variant, construction, re-rank, brute force and the exchange kernel on inputs small enough for a 50x slow-down.

    compute-sanitizer --tool memcheck python tools/sanitizer_cases.py [variant ...]

Variants: cta (default latency kernel), lat1 (one-warp latency variant), thr (throughput kernel, 24-warp plan),
dense (28-warp plan), fed (pageable batch fed to the running kernel), build, rerank, brute, exchange.
Each search result is compared with the golden output of the same call made earlier WITHOUT the sanitizer
(tests/golden), so the run is also a parity check.
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ENV = {"cta": {"FNB_LAT": "2", "FNB_PF2_MINB": "1"},  # (the two-hop prefetch also on the goldens' short lists) "lat1": {"FNB_LAT": "1"}, "thr": {"FNB_LAT": "0", "FNB_DENSE": "0"},
       "dense": {"FNB_LAT": "0", "FNB_DENSE": "1"}, "fed": {"FNB_LAT": "0"}}


def child(variant: str) -> None:
    import json

    import numpy as np

    import flatnav_b200
    from flatnav_b200.data_type import DataType
    gold = os.path.join(ROOT, "tests", "golden")
    cases = json.load(open(os.path.join(gold, "golden.json")))
    DT = {"f32": DataType.float32, "u8": DataType.uint8, "i8": DataType.int8}
    if variant in ("cta", "lat1", "thr", "dense"):
        for case in cases[:4]:
            g = np.load(os.path.join(gold, case["name"] + ".npz"))
            cls = flatnav_b200.index.index_class("l2" if case["metric"] == "l2" else "angular", DT[case["dtype"]])
            ix = cls.load_index(os.path.join(gold, case["name"] + ".idx"))
            K, ef = case["runs"][0]
            q = g["queries"][:64]
            d, l = ix.search(q, K, ef)
            dr = g[f"dist_k{K}_ef{ef}"][:64]
            assert float(np.max(np.abs(d - dr) / np.maximum(np.abs(dr), 1e-6))) <= 1e-5, case["name"]
            print(f"[{variant}] {case['name']}: {ix.kernel_signature(64, K, ef)} ok", flush=True)
    elif variant == "fed":
        case = cases[0]
        g = np.load(os.path.join(gold, case["name"] + ".npz"))
        ix = flatnav_b200.index.IndexL2Float.load_index(os.path.join(gold, case["name"] + ".idx"))
        K, ef = case["runs"][0]
        reps = (1 << 20) // g["queries"][0].nbytes // g["queries"].shape[0] + 2  # >= 1 MB of pageable queries
        q = np.ascontiguousarray(np.tile(g["queries"], (reps, 1)))
        d, l = ix.search(q, K, ef)
        n = g["queries"].shape[0]
        assert np.array_equal(d[:n], d[n:2 * n]) and np.array_equal(l[-n:], l[:n])
        print(f"[fed] {q.shape[0]} queries ({q.nbytes >> 10} KB pageable) ok", flush=True)
    elif variant == "build":
        from flatnav_b200 import synthetic
        data = synthetic.make("latent", 3000, 32)
        ix = flatnav_b200.index.create("l2", 32, 3000, 16)
        ix.add(data[:2500], 48)
        ix.add(data[2500:], 48)
        d, l = ix.search(data[:256], 1, 48)
        assert (l[:, 0] == np.arange(256)).mean() > 0.95
        print("[build] 3000 nodes in two add() calls ok", flush=True)
    elif variant in ("rerank", "brute"):
        case = cases[0]
        g = np.load(os.path.join(gold, case["name"] + ".npz"))
        ix = flatnav_b200.index.IndexL2Float.load_index(os.path.join(gold, case["name"] + ".idx"))
        q = g["queries"][:32]
        os.environ["FNB_BF_MODE"] = "exact"
        db, lb = ix.bruteforce(q, 10)
        if variant == "rerank":
            _, cand = ix.search(q, 64, 64)
            d, l = ix.rerank(q, cand, 10)
            assert np.all(np.diff(d, axis=1) >= 0)
        print(f"[{variant}] ok", flush=True)
    elif variant == "exchange":
        import ctypes as C

        import torch

        from flatnav_b200 import _capi
        case = cases[0]
        g = np.load(os.path.join(gold, case["name"] + ".npz"))
        ix = flatnav_b200.index.IndexL2Float.load_index(os.path.join(gold, case["name"] + ".idx"))
        ex = C.c_void_p()
        _capi.check(_capi.lib().fnb_exchange_create(0, 0, 1, 4096, 16, C.byref(ex)))
        q = torch.from_numpy(g["queries"][:64]).cuda()
        od = torch.empty((64, 10), dtype=torch.float32, device="cuda")
        ol = torch.empty((64, 10), dtype=torch.int32, device="cuda")
        for _ in range(3):
            _capi.check(_capi.lib().fnb_search_sharded(ix._h, ex, q.data_ptr(), 64, 10, 50, 100, od.data_ptr(), ol.data_ptr(), None))
        torch.cuda.synchronize()
        _capi.check(_capi.lib().fnb_exchange_status(ex))
        d, l = ix.search(g["queries"][:64], 10, 50)
        assert np.array_equal(ol.cpu().numpy(), l)
        _capi.lib().fnb_exchange_free(ex)
        print("[exchange] one-rank exchange + merge kernel ok", flush=True)
    else:
        raise SystemExit(f"unknown variant {variant}")


def main() -> None:
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
        return
    variants = sys.argv[1:] or ["cta", "lat1", "thr", "dense", "fed", "build", "rerank", "brute", "exchange"]
    rc = 0
    for v in variants:  # one process per variant: the environment knobs are read once per process
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", v], env=dict(os.environ, **ENV.get(v, {})))
        rc |= r.returncode
    sys.exit(rc)


if __name__ == "__main__":
    main()
