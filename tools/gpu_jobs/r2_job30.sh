# round 2, job 30: the shipped two-hop prefetch: GPU suite with it forced on every small batch, compute-sanitizer over
# the CTA latency kernel with it on, the latency protocol (reference beside it)
mkdir -p gpurun_out
( time FNB_PF2_MINB=1 timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/r2_pytest_gpu_pf2_forced.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu_pf2_forced.log | head -3
for tool in memcheck racecheck synccheck; do
  ( time timeout 200 compute-sanitizer --tool $tool --target-processes all --print-limit 40 python tools/sanitizer_cases.py cta ) > gpurun_out/r2san_cta_pf2_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^real" gpurun_out/r2san_cta_pf2_$tool.log | sort | uniq -c | head -6; grep -c "\] .* ok" gpurun_out/r2san_cta_pf2_$tool.log
done
timeout 300 python tools/latency.py cfg1 --q 2000 --builder gpu --efs 32,64,100,200 --out gpurun_out/r2_latency_cfg1.json > gpurun_out/r2_latency.log 2>&1
timeout 300 python tools/latency.py cfg1 --paper --q 1000 --builder gpu --out gpurun_out/r2_latency_cfg1_paper.json > gpurun_out/r2_latency_paper.log 2>&1
python - <<'P'
import json
for f in ("gpurun_out/r2_latency_cfg1.json", "gpurun_out/r2_latency_cfg1_paper.json"):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "missing", e); continue
    for r in d["rows"]:
        print(f.split("/")[-1], r["ef"], round(r["search_single"]["latency_p50"], 4), round(r["search_single"]["latency_p99"], 4),
              "ref", round(r["reference_1thread"]["latency_p50"], 4), round(r["reference_1thread"]["latency_p99"], 4))
P
