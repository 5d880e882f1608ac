# round 2, job 20: two-hop prefetch of the CTA latency kernel (default for batches of <= 64 queries) against the same
# kernel without it (FNB_PF2_MAXQ=0) on the same box
mkdir -p gpurun_out
timeout 150 python tools/sanitizer_cases.py cta > gpurun_out/r2j20_cta_plain.log 2>&1; rc=$?
echo "goldens rc=$rc"; tail -4 gpurun_out/r2j20_cta_plain.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python tools/latency.py cfg1 --q 1000 --builder gpu --efs 32,64,100,200 --no-ref --out gpurun_out/r2j20_latency_pf2.json > gpurun_out/r2j20_latency_pf2.log 2>&1; tail -4 gpurun_out/r2j20_latency_pf2.log | cut -c1-330
FNB_PF2_MAXQ=0 timeout 300 python tools/latency.py cfg1 --q 1000 --builder gpu --efs 32,64,100,200 --no-ref --out gpurun_out/r2j20_latency_nopf2.json > gpurun_out/r2j20_latency_nopf2.log 2>&1; tail -4 gpurun_out/r2j20_latency_nopf2.log | cut -c1-330
