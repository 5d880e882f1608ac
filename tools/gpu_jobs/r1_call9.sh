# random-gather ceiling of the memory system + per-query latency protocol
mkdir -p gpurun_out
( time timeout 300 tools/gather_peak.bin --rows-mb 4096 ) > gpurun_out/gather_peak_4g.jsonl 2> gpurun_out/gather_peak.err
( time timeout 300 tools/gather_peak.bin --rows-mb 640 ) > gpurun_out/gather_peak_640m.jsonl 2>> gpurun_out/gather_peak.err
cat gpurun_out/gather_peak_4g.jsonl | cut -c1-220
timeout 600 python tools/latency.py cfg1 --q 2000 --out gpurun_out/latency_cfg1.json > gpurun_out/latency.log 2>&1; tail -12 gpurun_out/latency.log
