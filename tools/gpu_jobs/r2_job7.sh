# round 2, job 7: the CTA-per-query latency kernel — suite with each kernel variant forced, latency protocol
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r2j7_pytest.log 2>&1; tail -4 gpurun_out/r2j7_pytest.log
( time FNB_LAT=2 timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_concurrency.py::test_byte_paths_agree_in_subprocesses ) > gpurun_out/r2j7_pytest_lat2.log 2>&1; tail -4 gpurun_out/r2j7_pytest_lat2.log
( time FNB_LAT=0 timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2j7_pytest_lat0.log 2>&1; tail -4 gpurun_out/r2j7_pytest_lat0.log
timeout 900 python tools/latency.py cfg1 --q 1000 --builder gpu --efs 32,100,200 --out gpurun_out/r2j7_latency_cfg1.json > gpurun_out/r2j7_latency.log 2>&1; tail -3 gpurun_out/r2j7_latency.log | cut -c1-700
timeout 900 python tools/latency.py cfg1 --paper --q 1000 --builder gpu --out gpurun_out/r2j7_latency_cfg1_paper.json > gpurun_out/r2j7_latency_paper.log 2>&1; tail -6 gpurun_out/r2j7_latency_paper.log | cut -c1-400
