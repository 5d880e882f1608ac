# round 2, job 13: CTA kernel with a merge warp
mkdir -p gpurun_out
( time FNB_LAT=2 timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_concurrency.py::test_byte_paths_agree_in_subprocesses ) > gpurun_out/r2j13_pytest_lat2.log 2>&1; tail -3 gpurun_out/r2j13_pytest_lat2.log
for tool in synccheck racecheck; do ( time timeout 600 compute-sanitizer --tool $tool --target-processes all --print-limit 10 python tools/sanitizer_cases.py cta ) > gpurun_out/r2j13_san_$tool.log 2>&1; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2j13_san_$tool.log; done
timeout 900 python tools/latency.py cfg1 --q 1000 --builder gpu --efs 32,100,200 --no-ref --out gpurun_out/r2j13_latency_cfg1.json > gpurun_out/r2j13_latency.log 2>&1; tail -3 gpurun_out/r2j13_latency.log | cut -c1-200
timeout 900 python tools/concurrency_probe.py --threads 16 --out gpurun_out/r2j13_concurrency.json > gpurun_out/r2j13_concurrency.log 2>&1; tail -1 gpurun_out/r2j13_concurrency.log
