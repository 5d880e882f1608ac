# round 2, job 5: self-cleaning launch slots + programmatic dependent launch (adjacent launches overlap their tails)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2j5_pytest.log 2>&1; tail -6 gpurun_out/r2j5_pytest.log
timeout 600 python tools/ab_probe.py --libs "prev=variants/libprev.so,base=flatnav_b200/libflatnav_b200.so,nopdl=flatnav_b200/libflatnav_b200.so@FNB_NO_PDL=1" --cases "cfg1,u8,cfg2" --out gpurun_out/r2j5_ab.json 2>&1 | tee gpurun_out/r2j5_ab.log | tail -20
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j5_bench.json 2> gpurun_out/r2j5_bench.err; tail -c 1500 gpurun_out/r2j5_bench.json; tail -5 gpurun_out/r2j5_bench.err
