# round 2, job 9: latency kernel with a dedicated driver warp + flag-based completion; concurrency; sanitizer re-run
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r2j9_pytest.log 2>&1; tail -4 gpurun_out/r2j9_pytest.log
timeout 900 python tools/latency.py cfg1 --q 1000 --builder gpu --efs 32,100,200 --out gpurun_out/r2j9_latency_cfg1.json > gpurun_out/r2j9_latency.log 2>&1; tail -3 gpurun_out/r2j9_latency.log | cut -c1-500
timeout 900 python tools/latency.py cfg1 --q 500 --builder gpu --efs 100 --time-kernels --no-ref --out gpurun_out/r2j9_latency_cfg1_timed.json > gpurun_out/r2j9_latency_timed.log 2>&1; tail -1 gpurun_out/r2j9_latency_timed.log | cut -c1-500
timeout 900 python tools/concurrency_probe.py --threads 1,4,16,64 --out gpurun_out/r2j9_concurrency.json > gpurun_out/r2j9_concurrency.log 2>&1; tail -5 gpurun_out/r2j9_concurrency.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2j9_bench.json 2> gpurun_out/r2j9_bench.err; tail -c 1200 gpurun_out/r2j9_bench.json
bash tools/gpu_jobs/r2_sanitizer.sh
