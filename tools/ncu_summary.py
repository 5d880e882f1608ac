#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/ncu_summary.py <tag> <launches.csv> <full.ncu-rep>

Writes profiles/<tag>_launches.md (every launch of the bench command with its share of device time),
profiles/<tag>_search_kernel_metrics.csv (selected metrics of the --set full capture of the traversal kernel),
profiles/<tag>_search_kernel_hotlines.md (top source lines by executed instructions / stall samples) and
profiles/roofline_traffic.json (dram bytes per launch, read by bench.py for roofline.traffic).
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("FNB_PROFILES_OUT") or os.path.join(ROOT, "profiles")  # on the GPU box: a directory under gpurun_out/
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launches(tag, path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui, gi, bi = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    agg = collections.OrderedDict()
    seq = []
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        ms = v / {"ns": 1e6, "us": 1e3, "ms": 1.0, "s": 1e-3}.get(r[ui], 1e6)
        name = r[ki].split("(")[0]
        seq.append((name, r[gi], r[bi], ms))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: every kernel launch of the profiled bench.py command (ncu --metrics gpu__time_duration.sum "
                "--clock-control none)\n\nPer-launch times are cold-cache and serialised under the profiler: compare shares.\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f}% |\n")
        f.write("\n## launch sequence\n\n| # | kernel | grid | block | ms |\n|---:|---|---|---|---:|\n")
        for i, (n, g, b, ms) in enumerate(seq):
            f.write(f"| {i} | `{n}` | {g} | {b} | {ms:.4f} |\n")


def full(tag, rep, name="search_kernel", traffic_json=True):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    hdr0 = hdr
    got = {}
    with open(os.path.join(OUT, f"{tag}_{name}_metrics.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "value", "unit"])
        w.writerow(["kernel", vals[hdr.index("Kernel Name")], ""])
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                w.writerow([h, v, u])
                got[h] = (v, u)

    def to_bytes(v, u):
        x = float(v.replace(",", ""))
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]

    traffic = to_bytes(*got["dram__bytes_read.sum"]) + to_bytes(*got["dram__bytes_write.sum"])
    if traffic_json:
        # profiles/roofline_traffic.json: one record per profiled shape ("headline" is the one bench.py reads), each
        # naming the kernel instantiation and the commit it was captured at; bench.py refuses a record of another kernel
        path = os.path.join(OUT, "roofline_traffic.json")
        try:
            allrec = json.load(open(path))
            if "dram_bytes_per_launch" in allrec:  # round-1 format
                allrec = {}
        except Exception:
            allrec = {}
        kname = vals[hdr.index("Kernel Name")]
        kname = "".join(kname.split("(")[0].replace("void ", "").replace("fnb::", "").split())
        allrec[traffic_json if isinstance(traffic_json, str) else "headline"] = {
            "kernel": kname, "dram_bytes_per_launch": traffic, "git_sha": os.environ.get("GIT_SHA", "unknown"),
            "workload": os.environ.get("NCU_WORKLOAD", ""), "kernel_us": got.get("gpu__time_duration.sum", ("", ""))[0],
            "l2_sector_hit_rate_pct": got.get("lts__t_sector_hit_rate.pct", ("", ""))[0],
            "sectors_per_request": (float(got["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"][0].replace(",", "")) /
                                    max(1.0, float(got["l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"][0].replace(",", "")))),
            "source": f"profiles/{tag}_{name}_metrics.csv (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full "
                      "--clock-control none capture)"}
        json.dump(allrec, open(path, "w"), indent=1)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hi = next(i for i, r in enumerate(rows[:10]) if "Source" in r)
    hdr = rows[hi]
    ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    lines = []
    for r in rows[hi + 1:]:
        if r and r[0].strip().isdigit():
            try:
                lines.append((int(r[0]), r[1].strip(), int(r[ii]), int(r[si])))
            except ValueError:
                pass
    ti, ts = sum(x[2] for x in lines) or 1, sum(x[3] for x in lines) or 1
    with open(os.path.join(OUT, f"{tag}_{name}_hotlines.md"), "w") as f:
        f.write(f"# {tag}: hottest source lines of {vals[hdr0.index('Kernel Name')][:90]} (ncu --set full --import-source on; -lineinfo)\n\n"
                f"total warp-instructions attributed: {ti}, stall samples: {ts}\n\n"
                "| line | % instructions | % stall samples | source |\n|---:|---:|---:|---|\n")
        by = 3 if "--by-samples" in sys.argv else 2  # latency-bound kernels: rank by stall samples instead
        for ln, s, i, sa in sorted(lines, key=lambda x: -x[by])[:40 if by == 3 else 30]:
            f.write(f"| {ln} | {100 * i / ti:.1f} | {100 * sa / ts:.1f} | `{s[:110]}` |\n")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    tag = sys.argv[1]
    name = sys.argv[sys.argv.index("--name") + 1] if "--name" in sys.argv else "search_kernel"
    if sys.argv[2] != "-":
        launches(tag, sys.argv[2])
    shape = sys.argv[sys.argv.index("--shape") + 1] if "--shape" in sys.argv else True
    if sys.argv[3] != "-":
        full(tag, sys.argv[3], name=name, traffic_json=False if "--no-traffic" in sys.argv else shape)
    print("wrote profiles for", tag)
