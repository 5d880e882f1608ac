# round 2: the CTA latency kernel under ncu with the finest warp-state sampling interval (one search_single launch);
# the per-instruction and per-line CSVs are analysed off the box (profiles/r2_lat_cta_kernel_hotlines.md,
# profiles/experiments/r2_latency_cta_kernel_warp_states.json)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f --warp-sampling-interval 0 --warp-sampling-buffer-size 536870912 -k regex:fnb_search_cta_kernel -c 1 -o /tmp/p_lat python tools/ncu_one.py cfg1 --single --launches 4 > gpurun_out/r2_ncu_lat.log 2>&1
tail -2 gpurun_out/r2_ncu_lat.log
ncu -i /tmp/p_lat.ncu-rep --page source --print-source sass --csv > gpurun_out/r2_lat_sass.csv 2> gpurun_out/r2_lat_sass.err
ncu -i /tmp/p_lat.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r2_lat_source.csv 2>> gpurun_out/r2_lat_sass.err
ncu -i /tmp/p_lat.ncu-rep --page raw --csv > gpurun_out/r2_lat_raw.csv 2>> gpurun_out/r2_lat_sass.err
ls -la gpurun_out/r2_lat_*.csv
