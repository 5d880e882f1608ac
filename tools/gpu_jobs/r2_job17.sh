# round 2, job 17: the concurrency record with longer runs (65536 queries per thread count)
mkdir -p gpurun_out
nproc > gpurun_out/r2j17_cpu.txt
timeout 600 python tools/concurrency_probe.py --q 65536 --threads 1,4,8,16,32,64,128 --out gpurun_out/r2j17_concurrency.json > gpurun_out/r2j17_concurrency.log 2>&1; tail -7 gpurun_out/r2j17_concurrency.log
FNB_NO_COMBINE=1 timeout 400 python tools/concurrency_probe.py --q 65536 --threads 8,16,32,64 --out gpurun_out/r2j17_concurrency_nocombine.json > gpurun_out/r2j17_concurrency_nocombine.log 2>&1; tail -4 gpurun_out/r2j17_concurrency_nocombine.log
