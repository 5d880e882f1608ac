#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the FlatNav search hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of queries: `Q` queries searched with
(K, ef_search, num_initializations=100) on the workload below.  One JSON line is printed by rank 0.

Workload (BASELINE.json configs[0], the configuration the north-star target is quoted on):
  synthetic 1M x 128 float32, squared L2, M=32, ef_construction=100, 10k queries per step, K=10,
  ef_search=100.  Data: the "latent" generator of flatnav_b200/synthetic.py (rank-16 Gaussian latent +
  0.1 noise): on the README's literal IID Gaussians recall@10 >= 0.95 is unreachable.
  The graph: the file the reference arm built (unmodified reference, cached) when it is there, so that both arms search
  the same graph; otherwise one built by this engine's GPU construction.  `index_build` says which.

Numbers:
  value     whole-job QPS with queries and outputs resident in HBM (kernel-only path, fnb_search_device),
            CUDA events around the timed steps, max over ranks.
  e2e       the same metric through the public host-buffer API (flatnav_b200 ... .search(numpy) ->
            fnb_search): pinned host queries -> H2D -> kernel -> D2H results, every step.
  e2e_pageable  the same call with what a reference-binding caller passes: ordinary (pageable) numpy in, fresh numpy out.
  e2e_two_callers  the e2e call made by two host threads on alternate steps, each on its own pinned buffers (the engine is
            re-entrant: one caller's launch fills the tail of the other's).  Reported beside e2e, never instead of it.
  roofline  HBM bound. achieved = algorithmic bytes per launch / mean kernel duration (per-launch CUDA
            events recorded inside the timed region); algorithmic bytes = n_dist*D*s + n_hops*M*4 + Q*D*s
            + Q*K*8 with n_dist / n_hops counted by the kernel.  peak = MEASURED_PEAKS.json
            hbm_gbs (else the 6650 GB/s fallback of B200_PROFILING.md).  traffic = dram bytes per launch from the
            committed ncu capture, but only when that capture is of the kernel instantiation this run launched
            (profiles/roofline_traffic.json names it), else null.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, executeInParallel over Index::search) on this host's
            cores, on a bounded sample of the same queries.
`config` is the static description of the workload and is identical in both arms; everything measured is a
top-level key.  Further top-level records (outside the K timed steps, each with its own CUDA-event timing):
  sustained       the headline step repeated for >= 2 s (thermal steady state; the K-step region is tens of ms)
  strong_scaling  BASELINE.json configs[2] shape: ONE 100k-query batch over a 10M x 96 float32 graph (GPU-built once,
                  loaded by every rank), split evenly across the ranks, results all-gathered (NCCL) — total queries
                  fixed as N grows
  sharded         BASELINE.json configs[4] shape: 12.5M x 128 uint8 per rank (GPU-built, labels = global ids), every
                  rank answers all queries on its shard, global top-K by the peer-memory exchange kernel and by the
                  NCCL all-gather + merge baseline (bit-identical), recall against the exact global ground truth
With --impl reference the whole line is that CPU implementation instead (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = dict(name="cfg1: synthetic 1Mx128 f32 L2 (latent r=16 sigma=0.1), M=32, efc=100, Q=10000/step, K=10, ef_search=100",
                gen="latent", n=1_000_000, dim=128, metric="l2", M=32, efc=100, Q=10_000, K=10, ef=100, ninit=100)
STRONG = dict(name="cfg3: synthetic 10Mx96 f32 L2 (latent), M=32, efc=100, one 100k-query batch split across the ranks, K=10, ef_search=100",
              gen="latent", n=10_000_000, dim=96, Q=100_000, K=10, ef=100, steps=10)
SHARDED = dict(name="cfg5: synthetic 12.5Mx128 uint8 L2 (latent-u8) per rank, M=32, efc=100, Q=10000, K=10, ef_search=100 per shard",
               gen="latent-u8", n_shard=12_500_000, dim=128, Q=10_000, K=10, ef=100, steps=20)
METRIC = "QPS at recall@10>=0.95"
RECALL_FLOOR = 0.95
N_QUERY_BATCHES = 8           # distinct query batches rotated over the steps
CPU_SAMPLE_Q = 4000           # queries per reference-arm step / cpu_baseline sample
SUSTAINED_SECONDS = 2.0


def env_int(name: str, default: int) -> int:
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def apply_env_overrides() -> None:
    """Development knobs (smaller N for quick runs); the driver runs with none set."""
    for key in ("n", "dim", "Q", "K", "ef", "M", "efc"):
        v = os.environ.get("FNB_BENCH_" + key.upper())
        if v:
            WORKLOAD[key] = int(v)
            WORKLOAD["name"] += f" [override {key}={v}]"
    b = os.environ.get("FNB_BENCH_BUILDER")
    if b:
        WORKLOAD["builder"] = b
        WORKLOAD["name"] += f" [override builder={b}]"
    g = os.environ.get("FNB_BENCH_GEN")
    if g:
        WORKLOAD["gen"] = g
        WORKLOAD["name"] += f" [override gen={g}]"
    for key, tgt, fld in (("FNB_BENCH_STRONG_N", STRONG, "n"), ("FNB_BENCH_STRONG_Q", STRONG, "Q"),
                          ("FNB_BENCH_SHARD_N", SHARDED, "n_shard"), ("FNB_BENCH_SHARD_Q", SHARDED, "Q")):
        v = os.environ.get(key)
        if v:
            tgt[fld] = int(v)
            tgt["name"] += f" [override {fld}={v}]"


def static_config(world: int) -> dict:
    """The workload description both arms print, key for key (nothing measured in here)."""
    w = WORKLOAD
    return {"workload": w["name"], "n": w["n"], "dim": w["dim"], "metric": w["metric"], "M": w["M"],
            "ef_construction": w["efc"], "queries_per_step": w["Q"], "K": w["K"], "ef_search": w["ef"],
            "num_initializations": w["ninit"], "generator": w["gen"],
            "l2_policy": "index (644 MB of vectors+links) >> 126 MB L2; a different query batch every step",
            "parallelism": f"replicated index, {world} rank(s) x {w['Q']} queries per step (query sharding)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows: list[list[str]] = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 9:
                self.rows.append(parts)

    def mark(self) -> int:
        return len(self.rows)

    def stop(self, lo: int = 0, hi: int | None = None) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        return self.summary(lo, hi)

    def summary(self, lo: int = 0, hi: int | None = None) -> dict:
        rows = self.rows[lo:hi] or self.rows
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(rows)}


def measured_peak() -> tuple[float, str]:
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def norm_sig(s: str) -> str:
    return "".join(str(s).split())


def profile_traffic(kernel_sig: str, shape: str = "headline") -> tuple[float | None, str]:
    """dram bytes per launch from the committed `ncu --set full` capture — accepted only if that capture profiled the
    kernel instantiation this run launches (tools/ncu_summary.py writes the name and the git sha next to the bytes)."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        rec = rec.get(shape, rec)
        if norm_sig(rec.get("kernel", "")) != norm_sig(kernel_sig):
            return None, f"profiles/roofline_traffic.json is of {rec.get('kernel')!r}, this run launched {kernel_sig!r}: refused"
        return float(rec["dram_bytes_per_launch"]), f"ncu capture {rec.get('source')} at {rec.get('git_sha')}"
    except Exception as e:  # no capture committed
        return None, f"no usable profiles/roofline_traffic.json ({type(e).__name__})"


def make_queries(n_batches: int) -> list[np.ndarray]:
    from flatnav_b200 import synthetic
    w = WORKLOAD
    allq = synthetic.make(w["gen"], w["Q"] * n_batches, w["dim"], queries=True)
    return [np.ascontiguousarray(allq[i * w["Q"]:(i + 1) * w["Q"]]) for i in range(n_batches)]


def recall_of(found: np.ndarray, truth: np.ndarray) -> float:
    K = truth.shape[1]
    return float(np.mean([len(set(a.tolist()) & set(b.tolist())) / K for a, b in zip(found, truth)]))


# ---------------------------------------------------------------------------------------------------
def run_reference(args, rank: int, world: int) -> None:
    """Reference arm: the unmodified reference's batched CPU search (oracle/_ref) on this host's cores."""
    if rank != 0:
        return
    from oracle import refbin
    from tools.workload import ensure_index
    w = WORKLOAD
    if not refbin.available():
        emit({"impl": "reference", "unavailable": "oracle/_ref reference binary cannot run on this host"})
        return
    path, build_info = ensure_index(w["gen"], w["n"], w["dim"], w["metric"], w["M"], w["efc"], builder=w.get("builder", "reference"))
    cores = os.cpu_count() or 1
    nq = min(w["Q"], CPU_SAMPLE_Q)
    queries = make_queries(1)[0][:nq]
    # one process invocation = `warmup` untimed passes + `steps` timed passes of the bounded sample (the harness itself
    # runs one untimed pass; the other warm-up passes are the first entries of qps_runs, dropped here)
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    _, _, info = refbin.search(path, w["metric"], queries, w["K"], w["ef"], w["ninit"], threads=cores,
                               reps=steps + warmup - 1, want_results=False)
    runs = info["qps_runs"][warmup - 1:]
    total_s = sum(nq / r for r in runs)
    qps = nq * len(runs) / total_s
    sample = (f"{nq} of {w['Q']} queries per step, {len(runs)} steps after {warmup} warm-up passes, {cores} threads, "
              f"{refbin.isa()} build")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": len(runs),
        "warmup": warmup, "ms_per_step": total_s / len(runs) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": static_config(max(1, args.gpus)),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "index_build": build_info,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------
_RESULT_FD = None


def claim_stdout() -> None:
    """stdout carries exactly ONE line, the result: everything else a library prints there from native code (NCCL's
    version banner under torchrun, for one) is sent to stderr for the length of the run."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def log(msg: str) -> None:
    sys.stderr.write(f"[bench {time.strftime('%H:%M:%S')}] {msg}\n")
    sys.stderr.flush()


class Ctx:
    """Process-group plumbing shared by the legs."""

    def __init__(self, rank, world, local):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.rank, self.world, self.local = torch, dist, rank, world, local

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals: float) -> list[float]:
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def all_ok(self, err: Exception | None, what: str) -> None:
        """Every rank reaches this point; if any rank failed `what`, all raise (so nobody waits in a collective)."""
        (bad,) = self.max_over_ranks(0.0 if err is None else 1.0)
        if err is not None:
            raise err
        if bad:
            raise RuntimeError(f"{what} failed on another rank")

    def sum_over_ranks(self, *vals: float) -> list[float]:
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]


def algo_bytes(info: dict, nd: int, nh: int, Q: int, K: int) -> int:
    return int(nd * info["data_size_bytes"] + nh * info["max_edges_per_node"] * 4 + Q * info["data_size_bytes"] + Q * K * 8)


# ---------------------------------------------------------------------------------------------------
def leg_strong(cx: Ctx) -> dict:
    """BASELINE.json configs[2]: one fixed batch split across the ranks of a replicated 10M x 96 graph (strong scaling)."""
    import flatnav_b200
    from flatnav_b200 import synthetic
    from flatnav_b200.distributed import QueryShardedSearcher, partition
    from tools.workload import CACHE
    torch, dist = cx.torch, cx.dist
    s = STRONG
    Q, K, ef, n, dim = s["Q"], s["K"], s["ef"], s["n"], s["dim"]
    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, f"bench_strong_{s['gen']}_n{n}_d{dim}_l2_M32_efc100_gpubuilt.idx")
    build = {}
    t0 = time.time()
    ix, err = None, None
    if cx.rank == 0:
        try:
            data = synthetic.make_device(s["gen"], n, dim).cpu().numpy()
            ix = flatnav_b200.index.create("l2", dim, n, 32)
            ix.add(data, 100)
            del data
            build = {"builder": "flatnav_b200 GPU construction",
                     "build_seconds": round(ix.last_build_stats["device_ms"] * 1e-3, 2)}
            if cx.world > 1:
                ix.save(path + ".tmp")
                os.replace(path + ".tmp", path)
        except Exception as e:  # noqa: BLE001
            err = e
    cx.all_ok(err, "building the strong-scaling index")
    if ix is None:
        try:
            ix = flatnav_b200.index.IndexL2Float.load_index(path, devices=[cx.local])
        except Exception as e:  # noqa: BLE001
            err = e
    cx.all_ok(err, "loading the strong-scaling index")
    if cx.rank == 0 and cx.world > 1:
        try:
            os.remove(path)
        except OSError:
            pass
    build["setup_seconds"] = round(time.time() - t0, 1)
    info = ix.info
    dq_all = synthetic.make_device(s["gen"], Q, dim, queries=True)  # the same batch on every rank (same seed)
    start, count = partition(Q, cx.world, cx.rank)
    per = (Q + cx.world - 1) // cx.world
    dq = dq_all[start:start + count].contiguous()
    od = torch.full((per, K), float("inf"), dtype=torch.float32, device="cuda")
    ol = torch.full((per, K), -1, dtype=torch.int32, device="cuda")
    gd = torch.empty((cx.world * per, K), dtype=torch.float32, device="cuda")
    gl = torch.empty((cx.world * per, K), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        ix.search_device(dq.data_ptr(), count, K, ef, 100, od.data_ptr(), ol.data_ptr(), stream)
        if cx.world > 1:  # the result gather: every rank ends up with the whole batch's results
            dist.all_gather_into_tensor(gd, od)
            dist.all_gather_into_tensor(gl, ol)

    for _ in range(3):
        step()
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(s["steps"]):
        step()
    e1.record()
    cx.barrier()
    (ms,) = cx.max_over_ranks(e0.elapsed_time(e1) / s["steps"])
    # the traversal alone (no gather) on this rank's slice: what the collective adds
    e0.record()
    for _ in range(s["steps"]):
        ix.search_device(dq.data_ptr(), count, K, ef, 100, od.data_ptr(), ol.data_ptr(), stream)
    e1.record()
    cx.barrier()
    nd, nh, _ = ix.device_totals()
    ms_local_mine = e0.elapsed_time(e1) / s["steps"]
    (ms_local,) = cx.max_over_ranks(ms_local_mine)
    # recall of this rank's slice against the exact scan, averaged over the batch
    lab = ol[:count].cpu().numpy()
    _, gt = ix.bruteforce(dq.cpu().numpy(), K)
    hits, tot = cx.sum_over_ranks(recall_of(lab, gt) * count, count)
    # the public host-side path: numpy in, gathered numpy out on every rank (QueryShardedSearcher.search)
    q_host = dq_all.cpu().numpy()
    if cx.world > 1:
        qs = QueryShardedSearcher(ix)
        qs.search(q_host, K, ef)
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            hd, hl = qs.search(q_host, K, ef)
        host_ms = (time.perf_counter() - t0) / 3 * 1e3
        same = bool(np.array_equal(hl[start:start + count], lab))
    else:
        ix.search(q_host, K, ef)
        t0 = time.perf_counter()
        for _ in range(3):
            hd, hl = ix.search(q_host, K, ef)
        host_ms = (time.perf_counter() - t0) / 3 * 1e3
        same = bool(np.array_equal(hl, lab))
    (host_ms,) = cx.max_over_ranks(host_ms)
    peak, _ = measured_peak()
    ab = algo_bytes(info, nd, nh, count, K)
    out = {"workload": s["name"], "n_gpus": cx.world, "scaling": "strong", "queries_total": Q, "queries_per_rank": per,
           "steps": s["steps"], "ms_per_batch": ms, "qps": Q / (ms * 1e-3), "ms_traversal_only": ms_local,
           "qps_traversal_only": Q / (ms_local * 1e-3), "gather": "nccl all_gather_into_tensor x2" if cx.world > 1 else "none",
           "recall_at_k": round(hits / tot, 4), "host_api_ms_per_batch": host_ms, "host_api_qps": Q / (host_ms * 1e-3),
           "host_api_equals_device_path": same,
           "roofline_frac_rank0": ab / (ms_local_mine * 1e-3) / 1e9 / peak, "index_build": build}
    del ix
    torch.cuda.empty_cache()
    return out


def leg_sharded(cx: Ctx) -> dict:
    """BASELINE.json configs[4]: one 12.5M x 128 uint8 sub-graph per rank; peer-memory exchange vs NCCL all-gather."""
    import flatnav_b200
    from flatnav_b200 import synthetic
    from flatnav_b200.data_type import DataType
    from flatnav_b200.distributed import DatasetShardedSearcher, merge_topk_cuda
    torch, dist = cx.torch, cx.dist
    s = SHARDED
    Q, K, ef, n_shard, dim = s["Q"], s["K"], s["ef"], s["n_shard"], s["dim"]
    t0 = time.time()
    err = None
    try:
        data = synthetic.make_device(s["gen"], n_shard, dim, stream=cx.rank + 1).cpu().numpy()
        ix = flatnav_b200.index.create("l2", dim, n_shard, 32, DataType.uint8)
        ix.add(data, 100, labels=np.arange(cx.rank * n_shard, (cx.rank + 1) * n_shard, dtype=np.int32))
        del data
    except Exception as e:  # noqa: BLE001
        err = e
    cx.all_ok(err, "building a shard")
    build = {"builder": "flatnav_b200 GPU construction", "build_seconds": round(ix.last_build_stats["device_ms"] * 1e-3, 2),
             "setup_seconds": round(time.time() - t0, 1)}
    info = ix.info
    dq = synthetic.make_device(s["gen"], Q, dim, queries=True)
    q_host = dq.cpu().numpy()
    stream = torch.cuda.current_stream().cuda_stream
    # exact global ground truth: per-shard exact scan (tensor-core filter + exact re-rank), gathered and merged
    gd_, gl_ = ix.bruteforce(q_host, K)
    td, tl = torch.from_numpy(gd_).cuda(), torch.from_numpy(gl_).cuda()
    if cx.world > 1:
        ad = torch.empty((cx.world * Q, K), dtype=torch.float32, device="cuda")
        al = torch.empty((cx.world * Q, K), dtype=torch.int32, device="cuda")
        dist.all_gather_into_tensor(ad, td)
        dist.all_gather_into_tensor(al, tl)
        _, gt = merge_topk_cuda(ad.view(cx.world, Q, K), al.view(cx.world, Q, K), K)
        gt = gt.cpu().numpy()
    else:
        gt = gl_

    def timed(fn):
        for _ in range(3):
            fn()
        cx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(s["steps"]):
            r = fn()
        e1.record()
        cx.barrier()
        (ms,) = cx.max_over_ranks(e0.elapsed_time(e1) / s["steps"])
        return ms, r

    od = torch.empty((Q, K), dtype=torch.float32, device="cuda")
    ol = torch.empty((Q, K), dtype=torch.int32, device="cuda")
    ms_local, _ = timed(lambda: ix.search_device(dq.data_ptr(), Q, K, ef, 100, od.data_ptr(), ol.data_ptr(), stream))
    nd, nh, _ = ix.device_totals()
    peak, _ = measured_peak()
    frac = algo_bytes(info, nd, nh, Q, K) / (ms_local * 1e-3) / 1e9 / peak
    out = {"workload": s["name"], "n_gpus": cx.world, "n_total": cx.world * n_shard, "n_shard": n_shard, "steps": s["steps"],
           "ms_traversal_only": ms_local, "shard_kernel_roofline_frac": frac, "shard_kernel_qps": Q / (ms_local * 1e-3),
           "index_build": build}
    if cx.world > 1:
        res = {}
        for mode in ("nccl", "peer"):
            sh = DatasetShardedSearcher(ix, exchange=mode, max_queries=Q, max_k=K)
            ms, (rd, rl) = timed(lambda: sh.search_tensors(dq, K, ef))
            sh.check_status()
            res[mode] = (ms, rd.cpu().numpy(), rl.cpu().numpy())
            sh.close()
        same = bool(np.array_equal(res["nccl"][1], res["peer"][1]) and np.array_equal(res["nccl"][2], res["peer"][2]))
        out.update({"exchange": "peer", "ms_per_step": res["peer"][0], "qps": Q / (res["peer"][0] * 1e-3),
                    "ms_exchange_merge": res["peer"][0] - ms_local,
                    "nccl_baseline": {"ms_per_step": res["nccl"][0], "qps": Q / (res["nccl"][0] * 1e-3),
                                      "ms_exchange_merge": res["nccl"][0] - ms_local,
                                      "collective": "all_gather_into_tensor x2 + fnb_merge_topk"},
                    "recall_at_k": round(recall_of(res["peer"][2], gt), 4), "peer_equals_nccl": same})
    else:
        out.update({"exchange": "none (one shard)", "ms_per_step": ms_local, "qps": Q / (ms_local * 1e-3),
                    "recall_at_k": round(recall_of(ol.cpu().numpy(), gt), 4)})
    del ix
    torch.cuda.empty_cache()
    return out


def guarded(name: str, fn, cx: Ctx):
    """A failing extra leg must not take the headline line with it (all ranks run the leg, so they fail together)."""
    t0 = time.time()
    try:
        r = fn(cx)
        r["leg_seconds"] = round(time.time() - t0, 1)
        log(f"{name}: done in {r['leg_seconds']} s")
        return r
    except Exception as e:  # noqa: BLE001
        log(f"{name}: FAILED {type(e).__name__}: {e}")
        return {"error": f"{type(e).__name__}: {e}"[:400]}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sustained / strong_scaling / sharded legs")
    args = ap.parse_args()
    apply_env_overrides()
    claim_stdout()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import flatnav_b200
    from tools.workload import cached_index, ensure_index

    w = WORKLOAD
    warmup = max(3, args.warmup)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; flatnav_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(minutes=6))
    cx = Ctx(rank, world, local)
    barrier = cx.barrier

    # ---- index: the file the reference arm built if it is cached (both arms then search the same graph), else one
    # built by this engine's own GPU construction (rank 0); loaded by every rank onto its own GPU.  Nothing under
    # oracle/ runs in this arm outside the cpu_baseline leg.
    def bench_index():
        if "builder" in w:
            return ensure_index(w["gen"], w["n"], w["dim"], w["metric"], w["M"], w["efc"], builder=w["builder"])
        hit = cached_index(w["gen"], w["n"], w["dim"], w["metric"], w["M"], w["efc"], builder="reference")
        return hit or ensure_index(w["gen"], w["n"], w["dim"], w["metric"], w["M"], w["efc"], builder="gpu")

    if rank == 0:
        path, build_info = bench_index()
    barrier()
    if rank != 0:
        path, build_info = bench_index()
    ix = flatnav_b200.index.IndexL2Float.load_index(path, devices=[local])
    info = ix.info

    # ---- inputs: distinct query batches (per rank: a rank-specific rotation of the same pool) ----
    batches = make_queries(N_QUERY_BATCHES)
    Q, K, ef, ninit = w["Q"], w["K"], w["ef"], w["ninit"]
    d_batches = [torch.from_numpy(b).cuda() for b in batches]
    pinned = [torch.from_numpy(b).pin_memory() for b in batches]
    d_dist = torch.empty((Q, K), dtype=torch.float32, device="cuda")
    d_lab = torch.empty((Q, K), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    # recall@K of the operating point, against exact ground truth computed on the GPU (outside the timed region)
    _, gt = ix.bruteforce(batches[0], K)
    _, lab0 = ix.search(batches[0], K, ef, ninit)
    recall = recall_of(lab0, gt)

    def step_device(i: int):
        b = d_batches[(i + rank) % N_QUERY_BATCHES]
        ix.search_device(b.data_ptr(), Q, K, ef, ninit, d_dist.data_ptr(), d_lab.data_ptr(), stream)

    # ---- kernel-only timing -------------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(warmup):
        step_device(i)
    barrier()
    # Nothing but the K launches goes on the stream between the two events: the engine's launches are adjacent (no
    # memset between them) and use programmatic dependent launch, so the CTAs of step i+1 fill the SMs that the tail of
    # step i leaves idle — an event record between steps would serialise them again.  The average launch duration of
    # the roofline is therefore (t_end - t_begin) / K.
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin.record()
    for i in range(args.steps):
        step_device(i)
    t_end.record()
    barrier()
    total_ms = t_begin.elapsed_time(t_end)
    kernel_ms = [total_ms / args.steps]
    # one isolated launch at a time (synchronised before and after): what a single 10k batch costs on an idle GPU
    single = []
    for i in range(7):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_device(i)
        b.record()
        torch.cuda.synchronize()
        single.append(a.elapsed_time(b))
    single_ms = float(statistics.median(single))
    # counters of the LAST step (every batch has the same size; counts vary by <1 % between batches)
    nd, nh, ns = ix.device_totals()
    ab = algo_bytes(info, nd, nh, Q, K)
    kernel_sig = ix.kernel_signature(Q, K, ef)

    # ---- sustained: the same step for >= SUSTAINED_SECONDS (the K-step region above lasts tens of milliseconds) ----
    sustained = None
    if not args.no_extras:
        n_sus = max(args.steps, int(SUSTAINED_SECONDS / max(total_ms / args.steps * 1e-3, 1e-6)))
        lo = sampler.mark()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        s0.record()
        for i in range(n_sus):
            step_device(i)
        s1.record()
        barrier()
        hi = sampler.mark()
        (sus_ms,) = cx.max_over_ranks(s0.elapsed_time(s1))
        sustained = {"steps": n_sus, "seconds": sus_ms * 1e-3, "qps": world * Q * n_sus / (sus_ms * 1e-3),
                     "ms_per_step": sus_ms / n_sus, "clocks": sampler.summary(lo, hi) if rank == 0 else None}
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end timing through the public host API (pinned host queries -> H2D -> kernel -> D2H) ----
    out_d = torch.empty((Q, K), dtype=torch.float32).pin_memory().numpy()
    out_l = torch.empty((Q, K), dtype=torch.int32).pin_memory().numpy()
    pinned_np = [p.numpy() for p in pinned]
    for i in range(warmup):
        ix.search(pinned_np[i % N_QUERY_BATCHES], K, ef, ninit, out=(out_d, out_l))
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        ix.search(pinned_np[(i + rank) % N_QUERY_BATCHES], K, ef, ninit, out=(out_d, out_l))
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    # two host threads, each making the same synchronous call on its own pinned buffers (steps split between them): the
    # engine is re-entrant, so the launch of one caller fills the tail of the other's — what a serving process with
    # more than one request thread sees.  Reported beside e2e, never instead of it.
    e2e_two_ms = None
    if not args.no_extras:
        import threading
        outs2 = [(torch.empty((Q, K), dtype=torch.float32).pin_memory().numpy(),
                  torch.empty((Q, K), dtype=torch.int32).pin_memory().numpy()) for _ in range(2)]
        gate = threading.Barrier(3)
        errs = []

        def caller(t: int, n: int):
            try:
                gate.wait()
                for i in range(t, n, 2):
                    ix.search(pinned_np[(i + rank) % N_QUERY_BATCHES], K, ef, ninit, out=outs2[t])
            except Exception as e:      # noqa: BLE001 — reported below, the record is dropped
                errs.append(repr(e))

        for n in (2 * warmup, args.steps):
            th = [threading.Thread(target=caller, args=(t, n)) for t in range(2)]
            for x in th:
                x.start()
            barrier()
            gate.wait()
            t0 = time.perf_counter()
            for x in th:
                x.join()
            e2e_two_ms = (time.perf_counter() - t0) * 1e3
            gate.reset()
        if errs:
            log(f"two-caller leg failed: {errs[0]}")
            e2e_two_ms = None
        barrier()
    # the same call as a reference-binding caller makes it: ordinary numpy in, fresh numpy arrays out
    for i in range(warmup):
        ix.search(batches[i % N_QUERY_BATCHES], K, ef, ninit)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        pd_, pl_ = ix.search(batches[(i + rank) % N_QUERY_BATCHES], K, ef, ninit)
    torch.cuda.synchronize()
    e2e_page_ms = (time.perf_counter() - t0) * 1e3
    barrier()

    # ---- max over ranks -----------------------------------------------------------------------------
    total_ms, e2e_ms, e2e_page_ms, mean_kernel_ms, single_ms, e2e_two_ms = cx.max_over_ranks(
        total_ms, e2e_ms, e2e_page_ms, float(statistics.mean(kernel_ms)), single_ms, e2e_two_ms or 0.0)
    del d_batches, pinned
    strong = sharded = None
    if not args.no_extras:
        strong = guarded("strong_scaling", leg_strong, cx)
        sharded = guarded("sharded", leg_sharded, cx)

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = ab / (mean_kernel_ms * 1e-3) / 1e9
        traffic, traffic_note = profile_traffic(kernel_sig)
        h2d, d2h = int(Q * info["data_size_bytes"]), int(Q * K * 8)
        line = {
            "metric": METRIC, "value": world * Q * args.steps / (total_ms * 1e-3), "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": static_config(world),
            "recall_at_k": round(recall, 4), "recall_ok": bool(recall >= RECALL_FLOOR), "index_build": build_info,
            "e2e": {"value": world * Q * args.steps / (e2e_ms * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "buffers": "pinned host queries and outputs (kernel reads / writes them in place)"},
            "e2e_pageable": {"value": world * Q * args.steps / (e2e_page_ms * 1e-3), "unit": "queries/s",
                             "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                             "buffers": "ordinary numpy in, fresh numpy arrays out (what the reference binding's callers pass)"},
            "e2e_two_callers": ({"value": world * Q * args.steps / (e2e_two_ms * 1e-3), "unit": "queries/s",
                                 "buffers": "two host threads, each calling search() on its own pinned buffers; "
                                            "the steps are split between them"} if e2e_two_ms else None),
            "gpu_launches": args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                         "kernel": kernel_sig, "algorithmic_bytes_per_launch": int(ab),
                         "n_dist_per_query": nd / Q, "n_hops_per_query": nh / Q, "kernel_ms_mean": mean_kernel_ms,
                         "launches": "adjacent on one stream, programmatic dependent launch (the next step's CTAs fill "
                                     "the tail of the previous one)",
                         "isolated_launch_ms": single_ms, "isolated_launch_qps": world * Q / (single_ms * 1e-3),
                         "isolated_launch_frac": ab / (single_ms * 1e-3) / 1e9 / peak,
                         "dram_side_frac": (traffic / (single_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         "note": "algorithmic bytes count every evaluated row, L2-served ones included (the 100 entry-probe "
                                 "rows every query reads, adjacency rows prefetched a hop earlier), so frac may exceed 1; "
                                 "dram_side_frac = traffic / isolated launch time / peak"},
            "sustained": sustained, "strong_scaling": strong, "sharded": sharded,
        }
        if not line["recall_ok"]:
            line["invalid"] = f"recall@{K} {recall:.4f} is below the {RECALL_FLOOR} the metric requires"
            log("WARNING: " + line["invalid"])
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(path, batches[0])
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(path: str, queries: np.ndarray) -> dict:
    from oracle import refbin
    w = WORKLOAD
    if not refbin.available():
        return {"value": None, "unit": "queries/s", "cores": 0, "kind": "reference",
                "sample": "oracle/_ref reference binary cannot run on this host"}
    cores = os.cpu_count() or 1
    nq = min(w["Q"], CPU_SAMPLE_Q)
    _, _, info = refbin.search(path, w["metric"], queries[:nq], w["K"], w["ef"], w["ninit"], threads=cores, reps=3,
                               want_results=False)
    return {"value": info["qps_best"], "unit": "queries/s", "cores": cores, "kind": "reference",
            "sample": f"{nq} of {w['Q']} queries, 1 warm-up + best of 3 passes, executeInParallel over Index::search "
                      f"with {cores} threads, reference compiled for {refbin.isa()}"}


if __name__ == "__main__":
    main()
