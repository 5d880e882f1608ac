# round 2, job 4: GPU suite + bench on the refactored host side (lanes, fed pageable path, rerank), A/B against round 1
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r2j4_pytest.log 2>&1; tail -15 gpurun_out/r2j4_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j4_bench.json 2> gpurun_out/r2j4_bench.err; tail -c 1500 gpurun_out/r2j4_bench.json; tail -5 gpurun_out/r2j4_bench.err
timeout 600 python tools/ab_probe.py --libs "prev=variants/libprev.so,base=flatnav_b200/libflatnav_b200.so" --cases "cfg1,u8,cfg1big,u8big,cfg2" --out gpurun_out/r2j4_ab.json 2>&1 | tee gpurun_out/r2j4_ab.log | tail -20
