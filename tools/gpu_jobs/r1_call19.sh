# merge with a 32-bit binary search + match.any rank fast path: A/B against the previous build, parity, latency
mkdir -p gpurun_out
L=flatnav_b200/libflatnav_b200.so
( time timeout 500 python tools/ab_probe.py --reps 5 --libs "prev=variants/libprev.so,new=$L" \
   --cases "${CASES:-cfg1,cfg2,u8,cfg1big,u8big}" --out gpurun_out/ab_merge.json ) > gpurun_out/ab_merge.log 2>&1
tail -16 gpurun_out/ab_merge.log
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log
( time FNB_LAT=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q ) > gpurun_out/pytest_gpu_lat.log 2>&1; grep -E "passed|failed|error" gpurun_out/pytest_gpu_lat.log
timeout 300 python tools/latency.py cfg1 --q 1000 --efs 32,100 --out gpurun_out/latency_cfg1_merge.json > gpurun_out/latency_merge.log 2>&1; tail -3 gpurun_out/latency_merge.log | cut -c1-330
