#!/usr/bin/env python
"""Generate tests/golden/: small index files BUILT BY THE UNMODIFIED REFERENCE plus the reference's own
search outputs on fixed queries.  Run in the build container (needs oracle/_ref, i.e. /root/reference):

    make -C oracle && python tools/make_golden.py

Each case <name> writes
    <name>.idx          the reference's cereal-serialised index (Index::saveIndex)
    <name>.npz          queries, and for every (K, ef): the reference's distances / labels
and the manifest golden.json lists the cases.  The reference has no golden search vectors of its own
(SURVEY.md §4), so these files are what pins the oracle port and the CUDA path on the GPU box, where
/root/reference does not exist.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flatnav_b200 import synthetic  # noqa: E402
from oracle import refbin  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name, metric, generator, N, D, M, efc, n_queries, [(K, ef), ...]
CASES = [
    ("l2_f32_d24", "l2", "latent", 2000, 24, 16, 64, 64, [(10, 50), (1, 16), (100, 100)]),
    ("ip_f32_d100", "ip", "latent-norm", 1200, 100, 8, 64, 64, [(10, 50), (5, 64)]),
    ("l2_u8_d32", "l2", "latent-u8", 2000, 32, 16, 64, 64, [(10, 50)]),
    ("ip_u8_d32", "ip", "latent-u8", 1500, 32, 16, 64, 64, [(10, 50)]),
    ("l2_i8_d32", "l2", "latent-i8", 2000, 32, 16, 64, 64, [(10, 50)]),
    ("ip_i8_d48", "ip", "latent-i8", 1500, 48, 16, 64, 64, [(10, 50)]),
    ("l2_f32_d7", "l2", "iid", 600, 7, 8, 32, 32, [(10, 40)]),       # residual (non multiple-of-4) dimension
    ("l2_f32_d200", "l2", "latent", 500, 200, 8, 32, 32, [(10, 40)]),  # > 32 chunks: whole-warp rows
    ("l2_f32_partial", "l2", "latent", 300, 16, 8, 32, 16, [(10, 30)]),  # cur_num_nodes < max_node_count (see below)
]


def main() -> None:
    if not refbin.available():
        sys.exit("oracle/_ref is not built: run `make -C oracle` where /root/reference exists")
    os.makedirs(OUT, exist_ok=True)
    manifest = []
    for name, metric, gen, n, d, M, efc, nq, runs in CASES:
        data = synthetic.make(gen, n, d)
        queries = synthetic.make(gen, nq, d, queries=True)
        idx = os.path.join(OUT, name + ".idx")
        refbin.build_index(data, metric, M, efc, idx, threads=1)  # 1 thread => deterministic graph
        if name.endswith("partial"):
            # emulate an index saved before it was full: bump max_node_count in the header and append
            # garbage nodes (the reference leaves them uninitialised)
            raw = bytearray(open(idx, "rb").read())
            node = int.from_bytes(raw[20:28], "little")
            extra = 57
            raw[28:36] = (n + extra).to_bytes(8, "little")
            raw += bytes(np.random.default_rng(7).integers(0, 256, node * extra, dtype=np.uint8))
            open(idx, "wb").write(raw)
        arrays = {"queries": queries}
        for K, ef in runs:
            dist, lab, _ = refbin.search(idx, metric, queries, K, ef, ninit=100, threads=1)
            arrays[f"dist_k{K}_ef{ef}"] = dist
            arrays[f"label_k{K}_ef{ef}"] = lab
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        manifest.append({"name": name, "metric": metric, "dtype": synthetic.dtype_code(data), "N": n, "D": d, "M": M,
                         "efc": efc, "runs": [list(r) for r in runs], "ref_isa": refbin.isa()})
        print(name, os.path.getsize(idx), "bytes")
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
