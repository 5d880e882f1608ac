#!/usr/bin/env python
"""Generate tests/golden/reorder.json and the Matrix Market fixture: outputs of the UNMODIFIED reference's
re-ordering (Index::doGraphReordering, Index.h:412-427) and link import (allocateNode + buildGraphLinks,
Index.h:187-272) on the golden index files.  Run in the build container after tools/make_golden.py:

    make -C oracle && python tools/make_golden_reorder.py

reorder.json: for every golden case and every strategy sequence, the SHA-256 of the file the reference saves after
re-ordering (whole file, and header + live nodes only: what lies beyond cur_num_nodes is uninitialised memory in the
reference and zeros in files written by fnb_index_save), and — as a cheap first diagnostic — the CRC32 of the permutation (old id -> new id; recovered from the
labels, which are 0..N-1 in node order in the golden files).
mtx_case.mtx / mtx_case.npz: a small edge list with the awkward cases (self-edges, more than M edges for a node,
nodes without edges, duplicate edges) and the SHA-256 of the index the reference builds from it.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import port, refbin  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SEQUENCES = [["gorder"], ["rcm"], ["gorder", "rcm"], ["rcm", "gorder"]]


def main() -> None:
    if not refbin.available():
        sys.exit("oracle/_ref is not built: run `make -C oracle` where /root/reference exists")
    cases = json.load(open(os.path.join(OUT, "golden.json")))
    doc = {"reorder": {}, "mtx": {}}
    with tempfile.TemporaryDirectory() as td:
        for c in cases:
            src = os.path.join(OUT, c["name"] + ".idx")
            entry = {}
            for seq in SEQUENCES:
                dst = os.path.join(td, "o.idx")
                refbin.reorder(src, c["metric"], c["dtype"], seq, dst)
                raw = open(dst, "rb").read()
                ix = port.OracleIndex(raw, port.L2)
                perm = np.empty(c["N"], dtype=np.uint32)
                perm[ix.labels()] = np.arange(c["N"], dtype=np.uint32)
                live = 60 + ix.node_size_bytes * ix.cur_num_nodes  # nodes beyond cur_num_nodes are uninitialised memory
                entry[",".join(seq)] = {"sha256": hashlib.sha256(raw).hexdigest(),
                                        "sha256_live": hashlib.sha256(raw[:live]).hexdigest(),
                                        "perm_crc32": zlib.crc32(perm.tobytes())}
            doc["reorder"][c["name"]] = entry
            print(c["name"], "ok")
        # ---- Matrix Market import ----
        n, d, M = 50, 8, 4
        rng = np.random.default_rng(11)
        data = rng.standard_normal((n, d)).astype(np.float32)
        edges = []
        for u in range(n):
            if u % 7 == 3:
                continue  # nodes without edges
            k = int(rng.integers(1, M + 3))  # up to M + 2 edges: the surplus must be dropped
            for v in rng.integers(0, n, k):
                edges.append((u, int(v)))
            if u % 5 == 0:
                edges.append((u, u))  # self-edge: leaves the slot available
                edges.append((u, (u + 1) % n))
        order = rng.permutation(len(edges))  # file order interleaves the source nodes
        edges = [edges[i] for i in order]
        mtx = os.path.join(OUT, "mtx_case.mtx")
        with open(mtx, "w") as f:
            f.write("%%MatrixMarket matrix coordinate pattern general\n% flatnav link import fixture\n")
            f.write(f"{n} {n} {M}\n")
            for u, v in edges:
                f.write(f"{u + 1} {v + 1}\n")
        dst = os.path.join(td, "m.idx")
        refbin.import_mtx(data, "l2", M, mtx, dst)
        raw = open(dst, "rb").read()
        np.savez_compressed(os.path.join(OUT, "mtx_case.npz"), data=data)
        doc["mtx"] = {"N": n, "D": d, "M": M, "sha256": hashlib.sha256(raw).hexdigest()}
    with open(os.path.join(OUT, "reorder.json"), "w") as f:
        json.dump(doc, f, indent=1)


if __name__ == "__main__":
    main()
