# round 2, job 16: combined launches for concurrent latency batches
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/r2j16_pytest.log 2>&1; tail -3 gpurun_out/r2j16_pytest.log
timeout 400 python tools/concurrency_probe.py --threads 1,4,8,16,32,64 --out gpurun_out/r2j16_concurrency.json > gpurun_out/r2j16_concurrency.log 2>&1; tail -6 gpurun_out/r2j16_concurrency.log
FNB_NO_COMBINE=1 timeout 400 python tools/concurrency_probe.py --threads 16,64 --out gpurun_out/r2j16_concurrency_nocombine.json > gpurun_out/r2j16_concurrency_nocombine.log 2>&1; tail -2 gpurun_out/r2j16_concurrency_nocombine.log
timeout 400 python tools/latency.py cfg1 --q 1000 --builder gpu --efs 100 --no-ref --out gpurun_out/r2j16_latency_cfg1.json > gpurun_out/r2j16_latency.log 2>&1; tail -1 gpurun_out/r2j16_latency.log | cut -c1-200
