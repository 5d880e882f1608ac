# round 2: compute-sanitizer over the small parity cases (tools/sanitizer_cases.py), one log per tool under gpurun_out/
# initcheck runs without the `fed` case: there the queries are written by copies on ANOTHER stream while the kernel
# runs (by design, ordered by the watermark), which the tool serialises behind the kernel — the launch then waits for
# its watchdog — and reports as reads of uninitialised memory.
mkdir -p gpurun_out
python tools/sanitizer_cases.py > gpurun_out/r2san_plain.log 2>&1; tail -3 gpurun_out/r2san_plain.log
for tool in memcheck racecheck synccheck initcheck; do
  cases=""; [ $tool = initcheck ] && cases="cta lat1 thr dense build rerank brute exchange"
  ( time timeout 900 compute-sanitizer --tool $tool --target-processes all --print-limit 40 python tools/sanitizer_cases.py $cases ) > gpurun_out/r2san_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^real" gpurun_out/r2san_$tool.log; grep -c "\] .* ok" gpurun_out/r2san_$tool.log
done
