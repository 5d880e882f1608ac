# round 2, last confirmation of the committed tree: the GPU suite, both bench arms (bench.py now also reports the
# two-caller end-to-end leg), smoke
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; cut -c1-300 gpurun_out/r2_bench_reference.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_b200.json 2> gpurun_out/r2_bench_b200.err; tail -c 400 gpurun_out/r2_bench_b200.json; tail -3 gpurun_out/r2_bench_b200.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
python - <<'P'
import json
d = json.load(open("gpurun_out/r2_bench_b200.json"))
print({k: (d[k]["value"] if isinstance(d.get(k), dict) and "value" in d[k] else d.get(k)) for k in ("value", "e2e", "e2e_two_callers", "e2e_pageable")})
P
