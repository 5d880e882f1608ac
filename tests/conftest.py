import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_cases():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


def golden_arrays(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_index_path(name):
    return os.path.join(GOLDEN, name + ".idx")


def rel_err(a, b):
    """max relative error with the 1e-5 parity tolerance's natural floor"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6))) if a.size else 0.0


def recall(found, truth):
    return float(np.mean([len(set(f.tolist()) & set(t.tolist())) / len(t) for f, t in zip(found, truth)]))


@pytest.fixture(scope="session")
def ref_cache(tmp_path_factory):
    """Directory for indexes built on the fly by the reference binary (oracle/_ref)."""
    return str(tmp_path_factory.mktemp("refidx"))


_BUILT = {}


def build_ref_index(cache_dir, metric, gen, n, d, M, efc):
    """Index built by the UNMODIFIED reference on this host (skips the test if the binary cannot run here)."""
    from flatnav_b200 import synthetic
    from oracle import refbin
    if not refbin.available():
        pytest.skip("oracle/_ref reference binary not runnable on this host")
    key = (metric, gen, n, d, M, efc)
    if key not in _BUILT:
        path = os.path.join(cache_dir, "_".join(map(str, key)) + ".idx")
        data = synthetic.make(gen, n, d)
        refbin.build_index(data, metric, M, efc, path, threads=max(1, min(8, os.cpu_count() or 1)))
        _BUILT[key] = path
    return _BUILT[key]
