# round 2, final: the records that go into profiles/ (suite with every kernel variant forced, bench, latency, concurrency)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
for v in 0 1 2; do ( time FNB_LAT=$v timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_concurrency.py::test_byte_paths_agree_in_subprocesses ) > gpurun_out/r2_pytest_gpu_lat${v}_forced.log 2>&1; tail -2 gpurun_out/r2_pytest_gpu_lat${v}_forced.log | head -1; done
( time FNB_LAT=0 FNB_DENSE=1 timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_concurrency.py::test_byte_paths_agree_in_subprocesses ) > gpurun_out/r2_pytest_gpu_dense_forced.log 2>&1; tail -2 gpurun_out/r2_pytest_gpu_dense_forced.log | head -1
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; cut -c1-300 gpurun_out/r2_bench_reference.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_b200.json 2> gpurun_out/r2_bench_b200.err; tail -c 600 gpurun_out/r2_bench_b200.json; tail -3 gpurun_out/r2_bench_b200.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
timeout 600 python tools/latency.py cfg1 --q 2000 --builder gpu --efs 32,64,100,200 --out gpurun_out/r2_latency_cfg1.json > gpurun_out/r2_latency.log 2>&1
timeout 900 python tools/latency.py cfg1 --paper --q 1000 --builder gpu --out gpurun_out/r2_latency_cfg1_paper.json > gpurun_out/r2_latency_paper.log 2>&1
timeout 600 python tools/concurrency_probe.py --threads 1,4,16,64 --out gpurun_out/r2_concurrency_cfg1.json > gpurun_out/r2_concurrency.log 2>&1; tail -4 gpurun_out/r2_concurrency.log
