# round 2, job 28: Python wrapper with cheaper pointer extraction: GPU suite + the latency protocol (both settings)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu.log | head -3
timeout 600 python tools/latency.py cfg1 --q 2000 --builder gpu --efs 32,64,100,200 --out gpurun_out/r2_latency_cfg1.json > gpurun_out/r2_latency.log 2>&1
timeout 900 python tools/latency.py cfg1 --paper --q 1000 --builder gpu --out gpurun_out/r2_latency_cfg1_paper.json > gpurun_out/r2_latency_paper.log 2>&1
python - <<'P'
import json
for f in ("gpurun_out/r2_latency_cfg1.json", "gpurun_out/r2_latency_cfg1_paper.json"):
    d = json.load(open(f))
    for r in d["rows"]:
        print(f.split("/")[-1], r["ef"], round(r["search_single"]["latency_p50"], 4), round(r["search_single"]["latency_p99"], 4),
              "ref", round(r["reference_1thread"]["latency_p50"], 4), round(r["reference_1thread"]["latency_p99"], 4))
P
