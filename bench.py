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
  the same graph; otherwise one built by this engine's GPU construction.  config.index_build says which.

Numbers:
  value     whole-job QPS with queries and outputs resident in HBM (kernel-only path, fnb_search_device),
            CUDA events around the timed steps, max over ranks.
  e2e       the same metric through the public host-buffer API (flatnav_b200 ... .search(numpy) ->
            fnb_search): pinned host queries -> H2D -> kernel -> D2H results, every step.
  roofline  HBM bound. achieved = algorithmic bytes per launch / mean kernel duration (per-launch CUDA
            events recorded inside the timed region); algorithmic bytes = n_dist*D*s + n_hops*M*4 + Q*D*s
            + Q*K*8 with n_dist / n_hops counted by the kernel.  peak = MEASURED_PEAKS.json
            hbm_gbs (else the 6650 GB/s fallback of B200_PROFILING.md).
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, executeInParallel over Index::search) on this host's
            cores, on a bounded sample of the same queries.
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
METRIC = "QPS at recall@10>=0.95"
N_QUERY_BATCHES = 8           # distinct query batches rotated over the steps
CPU_SAMPLE_Q = 4000           # queries per reference-arm step / cpu_baseline sample


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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def measured_peak() -> tuple[float, str]:
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profile_traffic() -> float | None:
    """dram bytes per launch of the traversal kernel from the committed ncu --set full capture, if any."""
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))["dram_bytes_per_launch"])
    except Exception:
        return None


def make_queries(n_batches: int) -> list[np.ndarray]:
    from flatnav_b200 import synthetic
    w = WORKLOAD
    allq = synthetic.make(w["gen"], w["Q"] * n_batches, w["dim"], queries=True)
    return [np.ascontiguousarray(allq[i * w["Q"]:(i + 1) * w["Q"]]) for i in range(n_batches)]


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
    path, _ = ensure_index(w["gen"], w["n"], w["dim"], w["metric"], w["M"], w["efc"], builder=w.get("builder", "reference"))
    cores = os.cpu_count() or 1
    nq = min(w["Q"], CPU_SAMPLE_Q)
    queries = make_queries(1)[0][:nq]
    # one process invocation = 1 warm-up pass + `steps` timed passes of the bounded sample
    reps = max(1, args.steps)
    _, _, info = refbin.search(path, w["metric"], queries, w["K"], w["ef"], w["ninit"], threads=cores, reps=reps,
                               want_results=False)
    runs = info["qps_runs"]
    total_s = sum(nq / r for r in runs)
    qps = nq * len(runs) / total_s
    sample = f"{nq} of {w['Q']} queries per step, {len(runs)} steps after 1 warm-up pass, {cores} threads, {refbin.isa()} build"
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": len(runs),
        "warmup": 1, "ms_per_step": total_s / len(runs) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": w["name"]},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
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


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    recall = float(np.mean([len(set(a.tolist()) & set(b.tolist())) / K for a, b in zip(lab0, gt)]))
    st0 = dict(ix.last_stats)

    def step_device(i: int):
        b = d_batches[(i + rank) % N_QUERY_BATCHES]
        ix.search_device(b.data_ptr(), Q, K, ef, ninit, d_dist.data_ptr(), d_lab.data_ptr(), stream)

    # ---- kernel-only timing -------------------------------------------------------------------------
    for i in range(warmup):
        step_device(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot_nd = tot_nh = 0
    t_begin.record()
    for i in range(args.steps):
        evs[i][0].record()
        step_device(i)
        evs[i][1].record()
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_begin.elapsed_time(t_end)
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    # counters of the LAST step (every batch has the same size; counts vary by <1 % between batches)
    nd, nh, ns = ix.device_totals()
    algo_bytes = nd * info["data_size_bytes"] + nh * info["max_edges_per_node"] * 4 + Q * info["data_size_bytes"] + Q * K * 8

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

    # ---- max over ranks -----------------------------------------------------------------------------
    t = torch.tensor([total_ms, e2e_ms, float(statistics.mean(kernel_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, mean_kernel_ms = [float(x) for x in t.tolist()]

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = algo_bytes / (mean_kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": world * Q * args.steps / (total_ms * 1e-3), "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "recall_at_k": round(recall, 4), "ef_search": ef,
                       "l2_policy": "index (644 MB of vectors+links) >> 126 MB L2; a different query batch every step",
                       "parallelism": f"replicated index, {world} rank(s) x {Q} queries per step (query sharding)",
                       "index_build": build_info},
            "e2e": {"value": world * Q * args.steps / (e2e_ms * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": int(Q * info["data_size_bytes"]), "d2h_bytes_per_step": int(Q * K * 8)},
            "gpu_launches": args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": profile_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(algo_bytes),
                         "n_dist_per_query": nd / Q, "n_hops_per_query": nh / Q, "kernel_ms_mean": mean_kernel_ms},
        }
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
