# round 2, job 8: profile of the CTA latency kernel, sanitizer pass, concurrent callers, re-ordering effect with L2 hit rates
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -k regex:fnb_search_cta_kernel -c 1 -o /tmp/p_lat python tools/ncu_one.py cfg1 --single --launches 4 > gpurun_out/r2j8_ncu_lat.log 2>&1
ncu -i /tmp/p_lat.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r2j8_lat_source.csv 2>/dev/null
ncu -i /tmp/p_lat.ncu-rep --page raw --csv > gpurun_out/r2j8_lat_raw.csv 2>/dev/null
timeout 900 python tools/concurrency_probe.py --threads 1,4,16,64 --out gpurun_out/r2j8_concurrency.json > gpurun_out/r2j8_concurrency.log 2>&1; tail -5 gpurun_out/r2j8_concurrency.log
timeout 900 ncu --metrics lts__t_sector_hit_rate.pct,dram__bytes_read.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:fnb_search_kernel --csv --log-file gpurun_out/r2j8_reorder_ncu.csv python tools/reorder_probe.py --n 1000000 --dim 128 --ef 32,100 --builder gpu --strategies "gorder;rcm" --profile --out gpurun_out/r2j8_reorder_cfg1.json > gpurun_out/r2j8_reorder.log 2>&1; tail -3 gpurun_out/r2j8_reorder.log
python tools/reorder_merge_ncu.py gpurun_out/r2j8_reorder_cfg1.json gpurun_out/r2j8_reorder_ncu.csv gpurun_out/r2j8_reorder_cfg1_l2.json
CASES="" bash tools/gpu_jobs/r2_sanitizer.sh
