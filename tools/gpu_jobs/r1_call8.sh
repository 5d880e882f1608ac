# one GPU call: A/B of kernel variants, dram bytes of the 400-byte-row config, bench both arms, the GPU test suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
( time timeout 480 python tools/ab_probe.py --reps 5 \
   --libs "base=flatnav_b200/libflatnav_b200.so,bulk=variants/libbulk.so,short7=variants/libshort7.so,short8=variants/libshort8.so,x2=variants/libx2.so,x2bulk=variants/libx2bulk.so" \
   --cases "${CASES:-cfg1,cfg2,u8,cfg4s}" --out gpurun_out/ab.json ) > gpurun_out/ab.log 2>&1
tail -25 gpurun_out/ab.log
P=$(ls data_cache/latent-norm_n1200000_d100_ip_M32_efc100_seed42_gpubuilt.idx)
timeout 120 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fnb_search_kernel -c 4 --csv --log-file gpurun_out/cfg2_dram.csv \
   python tools/ab_probe.py --libs x=x --reps 1 --child cfg2 $P > gpurun_out/cfg2_ncu.log 2>&1
tail -6 gpurun_out/cfg2_dram.csv
( time timeout 300 python bench.py ) > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; cat gpurun_out/bench_b200.json
( time timeout 300 python bench.py --impl reference --steps 10 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/pytest_gpu.log 2>&1
tail -30 gpurun_out/pytest_gpu.log
