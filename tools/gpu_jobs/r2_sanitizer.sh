# round 2: compute-sanitizer over the small parity cases (tools/sanitizer_cases.py), one log per tool under gpurun_out/
mkdir -p gpurun_out
python tools/sanitizer_cases.py > gpurun_out/r2san_plain.log 2>&1; tail -3 gpurun_out/r2san_plain.log
for tool in memcheck racecheck initcheck synccheck; do
  ( time timeout 1500 compute-sanitizer --tool $tool --target-processes all --print-limit 40 python tools/sanitizer_cases.py ${CASES:-} ) > gpurun_out/r2san_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|\] .* ok|real" gpurun_out/r2san_$tool.log | sort | uniq -c | sort -rn | head -30
done
