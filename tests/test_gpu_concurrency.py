"""GPU (-m gpu): fnb_search is re-entrant — concurrent callers get their own lane (stream, counters, staging) from the
replica's pool, like every thread of the reference gets its own visited set from VisitedSetPool
(include/flatnav/util/VisitedSetPool.h:154-172) under executeInParallel (util/Multithreading.h:18-48) — and the three
ways a host-buffer call moves its bytes (pinned in place, pageable fed while the kernel runs, plain staged copies)
return the same bytes."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import flatnav_b200
from conftest import ROOT, build_ref_index, golden_arrays, golden_cases, golden_index_path
from flatnav_b200 import synthetic

pytestmark = pytest.mark.gpu


def test_python_threads_search_concurrently(ref_cache):
    path = build_ref_index(ref_cache, "l2", "latent", 20000, 128, 32, 100)
    ix = flatnav_b200.index.IndexL2Float.load_index(path)
    q = synthetic.make("latent", 4096, 128, queries=True)
    want = ix.search(q, 10, 64)
    T, rounds = 12, 6
    got, errs = {}, []

    def work(t):
        try:
            for r in range(rounds):
                lo = (t * 331 + r * 97) % 3000
                n = (1, 7, 300, 1000)[(t + r) % 4]  # latency variant, small and large batches side by side
                d, l = ix.search(q[lo:lo + n], 10, 64)
                got[(t, r)] = (lo, n, d, l)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    assert len(got) == T * rounds
    for lo, n, d, l in got.values():
        np.testing.assert_array_equal(d.view(np.uint32), want[0][lo:lo + n].view(np.uint32))
        np.testing.assert_array_equal(l, want[1][lo:lo + n])


def test_fed_pageable_batch_equals_device_path(ref_cache):
    """>= 1 MB of pageable queries are fed to the running kernel chunk by chunk (watermark in device memory)"""
    import torch
    path = build_ref_index(ref_cache, "l2", "latent", 20000, 128, 32, 100)
    ix = flatnav_b200.index.IndexL2Float.load_index(path)
    q = synthetic.make("latent", 30000, 128, queries=True)  # 15 MB: 48 chunks
    d, l = ix.search(q, 10, 48)
    dq = torch.from_numpy(q).cuda()
    od = torch.empty((q.shape[0], 10), dtype=torch.float32, device="cuda")
    ol = torch.empty((q.shape[0], 10), dtype=torch.int32, device="cuda")
    ix.search_device(dq.data_ptr(), q.shape[0], 10, 48, 100, od.data_ptr(), ol.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d.view(np.uint32), od.cpu().numpy().view(np.uint32))
    np.testing.assert_array_equal(l, ol.cpu().numpy())
    nd, nh, ns = ix.device_totals()
    assert ix.last_stats["n_hops"] == nh and ns == 0
    # pinned buffers (used in place) and an unaligned pageable view give the same bytes
    qp = torch.from_numpy(q).pin_memory().numpy()
    out = (torch.empty((q.shape[0], 10), dtype=torch.float32).pin_memory().numpy(),
           torch.empty((q.shape[0], 10), dtype=torch.int32).pin_memory().numpy())
    ix.search(qp, 10, 48, out=out)
    np.testing.assert_array_equal(out[0].view(np.uint32), d.view(np.uint32))
    np.testing.assert_array_equal(out[1], l)


def test_byte_paths_agree_in_subprocesses(ref_cache, tmp_path):
    """FNB_NO_FEED / FNB_NO_STAGING / FNB_NO_ZEROCOPY force the other copy paths: same output file"""
    path = build_ref_index(ref_cache, "l2", "latent", 20000, 128, 32, 100)
    prog = ("import sys, numpy as np; sys.path.insert(0, %r); import flatnav_b200; from flatnav_b200 import synthetic;"
            "ix = flatnav_b200.index.IndexL2Float.load_index(%r); q = synthetic.make('latent', 9000, 128, queries=True);"
            "d, l = ix.search(q, 10, 40); d1, l1 = ix.search(q[:3], 10, 40);"
            "np.savez(sys.argv[1], d=d, l=l, d1=d1, l1=l1)") % (ROOT, path)
    outs = []
    for i, env in enumerate(({}, {"FNB_NO_FEED": "1"}, {"FNB_NO_FEED": "1", "FNB_NO_STAGING": "1"}, {"FNB_NO_ZEROCOPY": "1"})):
        out = str(tmp_path / f"o{i}.npz")
        subprocess.run([sys.executable, "-c", prog, out], check=True, env=dict(os.environ, **env))
        outs.append(np.load(out))
    for o in outs[1:]:
        for k in ("d", "l", "d1", "l1"):
            np.testing.assert_array_equal(o[k], outs[0][k])


def test_cpp_index_search_from_16_threads(tmp_path, ref_cache):
    """include/flatnav_b200/Index.h: Index::search fanned over 16 threads (the reference's executeInParallel pattern,
    bindings.cpp:196-212) returns what the serial loop returns; prints both throughputs"""
    path = build_ref_index(ref_cache, "l2", "latent", 20000, 128, 32, 100)
    q = synthetic.make("latent", 2048, 128, queries=True)
    exe = str(tmp_path / "concurrent_search")
    subprocess.run(["g++", "-std=c++17", "-O2", "-pthread", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "concurrent_search.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "flatnav_b200"), "-lflatnav_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "flatnav_b200")], check=True)
    qp = str(tmp_path / "q.bin")
    q.tofile(qp)
    r = subprocess.run([exe, path, qp, str(q.shape[0]), "10", "64", "16"], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "concurrent == serial" in r.stdout
    print(r.stdout)


def test_add_to_a_loaded_index_in_small_steps(tmp_path):
    """a loaded index holds exactly cur_num_nodes rows; add() makes room up to the header's max_node_count by itself,
    and many small add() calls reuse the construction scratch (reference-style incremental insertion)"""
    data = synthetic.make("latent", 3000, 32)
    ix = flatnav_b200.index.create("l2", 32, 3000, 16)
    ix.add(data[:2000], 64)
    p = str(tmp_path / "part.idx")
    ix.save(p)
    ld = flatnav_b200.index.IndexL2Float.load_index(p)
    assert ld.info["cur_num_nodes"] == 2000 and ld.info["max_node_count"] == 3000
    for lo in range(2000, 3000, 50):
        ld.add(data[lo:lo + 50], 64, labels=np.arange(lo, lo + 50))
    assert ld.info["cur_num_nodes"] == 3000
    with pytest.raises(ValueError):  # full (Index.h:356-361)
        ld.add(data[:1], 64)
    d, l = ld.search(data, 1, 64)
    assert (l[:, 0] == np.arange(3000)).mean() >= 0.97 and np.all(d[l[:, 0] == np.arange(3000), 0] == 0)
