# round 2, job 6: GPU suite on the PDL build; latency protocol baseline; cfg4 operating point; re-ordering effect (QPS + L2 hit rate)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r2j6_pytest.log 2>&1; tail -6 gpurun_out/r2j6_pytest.log
timeout 900 python tools/latency.py cfg1 --paper --q 1000 --builder gpu --out gpurun_out/r2j6_latency_cfg1_paper.json > gpurun_out/r2j6_latency.log 2>&1; tail -3 gpurun_out/r2j6_latency.log | cut -c1-600
timeout 900 python tools/sweep.py cfg4 --builder gpu --efs 100,200,300,400,512,768,1024,1536,2048 --iters 5 --ref-sample 500 --out gpurun_out/r2j6_sweep_cfg4.json > gpurun_out/r2j6_sweep_cfg4.log 2>&1; grep "\[sweep\] {" gpurun_out/r2j6_sweep_cfg4.log | cut -c1-330
timeout 900 python tools/reorder_probe.py --n 1000000 --dim 128 --ef 32,100 --builder gpu --strategies "gorder;rcm" --out gpurun_out/r2j6_reorder_cfg1.json > gpurun_out/r2j6_reorder_cfg1.log 2>&1; tail -6 gpurun_out/r2j6_reorder_cfg1.log
