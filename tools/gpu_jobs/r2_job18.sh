# round 2, job 18: CTA latency kernel with speculative row evaluation (FNB_LAT=3, the new default) against the plain
# CTA kernel (FNB_LAT=2) on the same box: goldens, latency A/B, the GPU suite with the variant forced, synccheck
mkdir -p gpurun_out
FNB_LAT=3 timeout 150 python tools/sanitizer_cases.py spec > gpurun_out/r2j18_spec_plain.log 2>&1; rc=$?
echo "goldens rc=$rc"; tail -5 gpurun_out/r2j18_spec_plain.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python tools/latency.py cfg1 --q 1000 --builder gpu --efs 32,64,100,200 --no-ref --out gpurun_out/r2j18_latency_spec.json > gpurun_out/r2j18_latency_spec.log 2>&1; tail -4 gpurun_out/r2j18_latency_spec.log | cut -c1-400
FNB_LAT=2 timeout 300 python tools/latency.py cfg1 --q 1000 --builder gpu --efs 32,64,100,200 --no-ref --out gpurun_out/r2j18_latency_cta.json > gpurun_out/r2j18_latency_cta.log 2>&1; tail -4 gpurun_out/r2j18_latency_cta.log | cut -c1-400
( time FNB_LAT=3 timeout 400 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2j18_pytest_lat3.log 2>&1; tail -4 gpurun_out/r2j18_pytest_lat3.log
( time timeout 300 compute-sanitizer --tool synccheck --target-processes all --print-limit 20 python tools/sanitizer_cases.py spec ) > gpurun_out/r2j18_synccheck.log 2>&1
grep -E "ERROR SUMMARY|^real| ok$" gpurun_out/r2j18_synccheck.log | tail -8
