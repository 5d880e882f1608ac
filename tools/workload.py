"""Benchmark / test infrastructure (NOT part of the product package): materialise a synthetic dataset and an
index file for it, cached under data_cache/ keyed by every parameter that shapes it.  Two builders write the same
file format: the unmodified reference (oracle/_ref/ref_flatnav_*: Index::addBatch + saveIndex) and this engine's own
GPU construction (csrc/build.cu).  `cached_index` only looks: the product arm of bench.py uses it to search the very
file the reference arm built when that is there, and builds its own otherwise, without running anything under oracle/."""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CACHE = os.environ.get("FNB_DATA_CACHE", os.path.join(ROOT, "data_cache"))


def index_name(gen: str, n: int, dim: int, metric: str, M: int, efc: int, rank: int = 16) -> str:
    r = "" if rank == 16 else f"_r{rank}"
    return f"{gen}{r}_n{n}_d{dim}_{metric}_M{M}_efc{efc}_seed42"


def cached_index(gen: str, n: int, dim: int, metric: str, M: int = 32, efc: int = 100, rank: int = 16, builder: str = "reference"):
    """(path, info) of an index file already in the cache, else None.  Never builds, never imports oracle/."""
    path = os.path.join(CACHE, index_name(gen, n, dim, metric, M, efc, rank) + ("_gpubuilt" if builder == "gpu" else "") + ".idx")
    meta = path + ".json"
    if os.path.exists(path) and os.path.exists(meta):
        info = json.load(open(meta))
        info["cached"] = True
        return path, info
    return None


def ensure_index(gen: str, n: int, dim: int, metric: str, M: int = 32, efc: int = 100, threads: int = 0,
                 rank: int = 16, builder: str = "reference"):
    """Return (path, info).  Builds the file if it is not cached: with the reference binary (default), or with this
    engine's GPU construction (builder="gpu": the large configs, where the CPU build takes minutes) — either way the
    file is in the reference's format and both arms of a benchmark read the same graph."""
    from flatnav_b200 import synthetic
    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, index_name(gen, n, dim, metric, M, efc, rank) + ("_gpubuilt" if builder == "gpu" else "") + ".idx")
    meta = path + ".json"
    if os.path.exists(path) and os.path.exists(meta):
        info = json.load(open(meta))
        info["cached"] = True
        return path, info
    if builder == "gpu":
        import numpy as np

        import flatnav_b200
        from flatnav_b200.data_type import DataType
        t0 = time.time()
        data = synthetic.make(gen, n, dim, rank=rank)
        t_gen = time.time() - t0
        dt = {"float32": DataType.float32, "uint8": DataType.uint8, "int8": DataType.int8}[data.dtype.name]
        t0 = time.time()
        ix = flatnav_b200.index.create("l2" if metric == "l2" else "angular", dim, n, M, dt)
        ix.add(data, efc)
        st = dict(ix.last_build_stats)
        ix.save(path + ".tmp")
        os.replace(path + ".tmp", path)
        info = {"build_seconds": round(st["device_ms"] * 1e-3, 3), "build_and_save_seconds": round(time.time() - t0, 2),
                "gen_seconds": round(t_gen, 2), "builder": "flatnav_b200 GPU construction", "cached": False}
        json.dump(info, open(meta, "w"))
        return path, info
    from oracle import refbin  # only the reference builder touches oracle/
    if not refbin.available():
        raise RuntimeError("no cached index and the reference builder (oracle/_ref) cannot run on this host")
    t0 = time.time()
    data = synthetic.make(gen, n, dim, rank=rank)
    t_gen = time.time() - t0
    info = refbin.build_index(data, metric, M, efc, path + ".tmp", threads=threads or (os.cpu_count() or 1))
    os.replace(path + ".tmp", path)
    info = {"build_seconds": info["seconds"], "build_threads": info["threads"], "gen_seconds": round(t_gen, 2),
            "builder": "reference " + refbin.isa(), "cached": False}
    json.dump(info, open(meta, "w"))
    return path, info
