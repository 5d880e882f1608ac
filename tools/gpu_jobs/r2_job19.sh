# round 2, job 19: where the speculating latency kernel's hop goes (debug build with counters)
mkdir -p gpurun_out
FNB_LIB_PATH=variants/libfnb_specdbg.so timeout 300 python tools/spec_probe.py --ef 100 --q 300 --out gpurun_out/r2j19_spec_probe.json > gpurun_out/r2j19_spec_probe.log 2>&1
tail -3 gpurun_out/r2j19_spec_probe.log
