# round 2, job 10: synccheck of the CTA kernel, concurrent callers, the paper's latency protocol, ef sweeps of every config
mkdir -p gpurun_out
( time timeout 600 compute-sanitizer --tool synccheck --target-processes all --print-limit 10 python tools/sanitizer_cases.py cta lat1 thr dense fed ) > gpurun_out/r2san_synccheck.log 2>&1; grep -E "ERROR SUMMARY" gpurun_out/r2san_synccheck.log
timeout 900 python tools/concurrency_probe.py --threads 1,4,16,64 --out gpurun_out/r2j10_concurrency.json > gpurun_out/r2j10_concurrency.log 2>&1; tail -4 gpurun_out/r2j10_concurrency.log
timeout 900 python tools/latency.py cfg1 --paper --q 1000 --builder gpu --out gpurun_out/r2j10_latency_cfg1_paper.json > gpurun_out/r2j10_latency_paper.log 2>&1; tail -6 gpurun_out/r2j10_latency_paper.log | cut -c1-300
timeout 900 python tools/latency.py cfg1 --q 2000 --builder gpu --efs 32,64,100,200 --out gpurun_out/r2j10_latency_cfg1.json > gpurun_out/r2j10_latency.log 2>&1
for c in cfg1 cfg2; do timeout 900 python tools/sweep.py $c --builder gpu --out gpurun_out/r2_sweep_$c.json > gpurun_out/r2_sweep_$c.log 2>&1; grep "\[sweep\] {" gpurun_out/r2_sweep_$c.log | cut -c1-200; done
timeout 1200 python tools/sweep.py cfg5shard --builder gpu --ref-sample 1000 --out gpurun_out/r2_sweep_cfg5shard.json > gpurun_out/r2_sweep_cfg5shard.log 2>&1; grep "\[sweep\] {" gpurun_out/r2_sweep_cfg5shard.log | cut -c1-200
timeout 1200 python tools/sweep.py cfg3 --builder gpu --ref-sample 1000 --out gpurun_out/r2_sweep_cfg3.json > gpurun_out/r2_sweep_cfg3.log 2>&1; grep "\[sweep\] {" gpurun_out/r2_sweep_cfg3.log | cut -c1-200
