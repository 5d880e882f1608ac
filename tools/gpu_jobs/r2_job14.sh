# round 2, job 14: CTA kernel with a merge warp (barrier counts fixed) — every step under a short timeout
mkdir -p gpurun_out
( time FNB_LAT=2 timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r2j14_pytest_lat2.log 2>&1; tail -3 gpurun_out/r2j14_pytest_lat2.log
if grep -q " passed" gpurun_out/r2j14_pytest_lat2.log && ! grep -q "failed\|Killed\|Terminated" gpurun_out/r2j14_pytest_lat2.log; then
  ( time FNB_LAT=2 timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_concurrency.py::test_byte_paths_agree_in_subprocesses ) > gpurun_out/r2j14_pytest_lat2_all.log 2>&1; tail -3 gpurun_out/r2j14_pytest_lat2_all.log
  for tool in synccheck racecheck; do ( time timeout 300 compute-sanitizer --tool $tool --target-processes all --print-limit 10 python tools/sanitizer_cases.py cta ) > gpurun_out/r2j14_san_$tool.log 2>&1; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2j14_san_$tool.log; done
  timeout 400 python tools/latency.py cfg1 --q 1000 --builder gpu --efs 32,100,200 --no-ref --out gpurun_out/r2j14_latency_cfg1.json > gpurun_out/r2j14_latency.log 2>&1; tail -3 gpurun_out/r2j14_latency.log | cut -c1-120
  timeout 400 python tools/concurrency_probe.py --threads 16 --out gpurun_out/r2j14_concurrency.json > gpurun_out/r2j14_concurrency.log 2>&1; tail -1 gpurun_out/r2j14_concurrency.log
fi
