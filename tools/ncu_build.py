#!/usr/bin/env python
"""Profiling target: a 1M x 128 float32 GPU construction (csrc/build.cu) for `ncu -k regex:build_..._kernel -s N -c 1`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flatnav_b200  # noqa: E402
from flatnav_b200 import synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
data = synthetic.make("latent", n, 128)
ix = flatnav_b200.index.create("l2", 128, n, 32)
ix.add(data, 100)
print("built", ix.last_build_stats, flush=True)
