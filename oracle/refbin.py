"""TEST INFRASTRUCTURE — run the UNMODIFIED reference compiled into oracle/_ref/ (oracle/Makefile).

The binaries are built in the build container (where /root/reference exists) and travel to the GPU
box with the repository snapshot; nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
DT = {np.dtype(np.float32): "f32", np.dtype(np.uint8): "u8", np.dtype(np.int8): "i8"}


def _cpu_flags() -> set:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def binary() -> str | None:
    """Path of the reference binary this host can execute, or None."""
    flags = _cpu_flags()
    v4 = {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= flags
    order = ["ref_flatnav_v4", "ref_flatnav_v3"] if v4 else ["ref_flatnav_v3"]
    if "avx2" not in flags:
        order = []
    for name in order:
        p = os.path.join(REF_DIR, name)
        if os.path.exists(p) and os.access(p, os.X_OK):
            return p
    return None


def available() -> bool:
    return binary() is not None


def isa() -> str:
    b = binary()
    return "none" if b is None else ("avx512" if b.endswith("v4") else "avx2")


def _run(args: list[str], timeout: float | None = None) -> dict:
    b = binary()
    if b is None:
        raise RuntimeError("oracle/_ref reference binary is not available on this host")
    out = subprocess.run([b] + args, check=True, capture_output=True, text=True, timeout=timeout)
    last = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")][-1]
    return json.loads(last)


def build_index(data: np.ndarray, metric: str, M: int, ef_construction: int, out_path: str, threads: int = 0,
                timeout: float | None = None, first_label: int = 0) -> dict:
    """Index::addBatch + saveIndex with the reference itself.  `metric` is "l2" or "ip".  Labels are
    first_label, first_label + 1, ... in row order (dataset shards carry global ids)."""
    data = np.ascontiguousarray(data)
    threads = threads or (os.cpu_count() or 1)
    with tempfile.NamedTemporaryFile(suffix=".bin", dir=os.path.dirname(os.path.abspath(out_path))) as f:
        data.tofile(f.name)
        return _run(["build", metric, DT[data.dtype], f.name, str(data.shape[0]), str(data.shape[1]), str(M),
                     str(ef_construction), str(threads), out_path, str(first_label)], timeout=timeout)


def search(index_path: str, metric: str, queries: np.ndarray, K: int, ef: int, ninit: int = 100, threads: int = 1,
           reps: int = 1, want_results: bool = True, timeout: float | None = None):
    """The reference's batched search loop (bindings.cpp:196-212).  Returns (dist, label, info)."""
    queries = np.ascontiguousarray(queries)
    Q = queries.shape[0]
    with tempfile.TemporaryDirectory() as td:
        qp = os.path.join(td, "q.bin")
        queries.tofile(qp)
        prefix = os.path.join(td, "out") if want_results else "-"
        info = _run(["search", metric, DT[queries.dtype], index_path, qp, str(Q), str(K), str(ef), str(ninit),
                     str(threads), str(reps), prefix], timeout=timeout)
        if not want_results:
            return None, None, info
        d = np.fromfile(prefix + ".dist.bin", dtype=np.float32).reshape(Q, K)
        l = np.fromfile(prefix + ".label.bin", dtype=np.int32).reshape(Q, K)
    return d, l, info


def latency(index_path: str, metric: str, queries: np.ndarray, K: int, ef: int, ninit: int = 100,
            timeout: float | None = None) -> np.ndarray:
    """Per-query wall-clock seconds of Index::search on one thread (the search_single loop of
    experiments/run-benchmark.py:66-84), measured inside the reference process after one warm-up pass."""
    queries = np.ascontiguousarray(queries)
    Q = queries.shape[0]
    with tempfile.TemporaryDirectory() as td:
        qp, op = os.path.join(td, "q.bin"), os.path.join(td, "lat.bin")
        queries.tofile(qp)
        _run(["latency", metric, DT[queries.dtype], index_path, qp, str(Q), str(K), str(ef), str(ninit), op],
             timeout=timeout)
        return np.fromfile(op, dtype=np.float64)


def reorder(index_path: str, metric: str, dtype: str, strategies, out_path: str, timeout: float | None = None) -> dict:
    """Index::loadIndex + doGraphReordering(strategies) + saveIndex with the reference itself (Index.h:412-427).
    `dtype` is "f32", "u8" or "i8"."""
    return _run(["reorder", metric, dtype, index_path, out_path, ",".join(strategies)], timeout=timeout)


def import_mtx(data: np.ndarray, metric: str, M: int, mtx_path: str, out_path: str, timeout: float | None = None) -> dict:
    """allocateNode for every row + buildGraphLinks(mtx) + saveIndex with the reference itself (Index.h:187-272)."""
    data = np.ascontiguousarray(data)
    with tempfile.NamedTemporaryFile(suffix=".bin", dir=os.path.dirname(os.path.abspath(out_path))) as f:
        data.tofile(f.name)
        return _run(["mtx", metric, DT[data.dtype], f.name, str(data.shape[0]), str(data.shape[1]), str(M), mtx_path,
                     out_path], timeout=timeout)
