mkdir -p gpurun_out
( FNB_LAT=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q ) > gpurun_out/pytest_lat_forced.log 2>&1; tail -3 gpurun_out/pytest_lat_forced.log
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_build.py -m gpu -x -q ) > gpurun_out/pytest_lat_auto.log 2>&1; tail -3 gpurun_out/pytest_lat_auto.log
timeout 600 python tools/latency.py cfg1 --q 1000 --efs ${EFS:-100} --out gpurun_out/latency_cfg1_lat.json > gpurun_out/latency_lat.log 2>&1; tail -3 gpurun_out/latency_lat.log
