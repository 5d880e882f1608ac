# round 2, job 15: flag wait that yields after ~10 us: concurrent callers and single-caller latency
mkdir -p gpurun_out
nproc > gpurun_out/r2j15_cpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/r2j15_cpu.txt
timeout 600 python tools/concurrency_probe.py --threads 1,4,8,16,32,64 --out gpurun_out/r2j15_concurrency.json > gpurun_out/r2j15_concurrency.log 2>&1; tail -6 gpurun_out/r2j15_concurrency.log
timeout 400 python tools/latency.py cfg1 --q 1000 --builder gpu --efs 100 --no-ref --out gpurun_out/r2j15_latency_cfg1.json > gpurun_out/r2j15_latency.log 2>&1; tail -1 gpurun_out/r2j15_latency.log | cut -c1-200
