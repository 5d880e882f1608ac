# round 2: ncu captures of the SHIPPED kernels (run with GIT_SHA=<sha> in the environment).  Summaries are written on the
# box (the .ncu-rep files are ~20 MB each, gpurun_out/ is capped at 64 MiB) into gpurun_out/profiles_r2/, then copied to
# profiles/ by hand.
mkdir -p gpurun_out/profiles_r2
export FNB_PROFILES_OUT=$PWD/gpurun_out/profiles_r2
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
# 1. the launch list of the bench command (headline path)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file /tmp/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2ncu_launches.log 2>&1
# 2. headline kernel: one isolated 10k-query launch of cfg1 (24-warp plan)
NCU_WORKLOAD="cfg1 ef=100 Q=10000, one isolated launch" timeout 600 $NCU -k regex:fnb_search_kernel -c 1 -o /tmp/p_cfg1 python tools/ncu_one.py cfg1 > gpurun_out/r2ncu_cfg1.log 2>&1
python tools/ncu_summary.py r2_cfg1 /tmp/launches.csv /tmp/p_cfg1.ncu-rep --shape headline
# 3. dense plan: 100k-query batch
NCU_WORKLOAD="cfg1 ef=100 Q=100000 (28-warp plan)" timeout 600 $NCU -k regex:fnb_search_kernel -c 1 -o /tmp/p_cfg1big python tools/ncu_one.py cfg1big > gpurun_out/r2ncu_cfg1big.log 2>&1
python tools/ncu_summary.py r2_cfg1_dense - /tmp/p_cfg1big.ncu-rep --shape cfg1_dense
# 4. uint8 128-byte rows (cfg5 shard shape, 4M nodes)
NCU_WORKLOAD="u8 4Mx128 ef=100 Q=10000" timeout 600 $NCU -k regex:fnb_search_kernel -c 1 -o /tmp/p_u8 python tools/ncu_one.py u8 > gpurun_out/r2ncu_u8.log 2>&1
python tools/ncu_summary.py r2_u8 - /tmp/p_u8.ncu-rep --shape u8
# 5. cfg2 (400-byte rows, inner product) and cfg4-shaped long rows
NCU_WORKLOAD="cfg2 1.2Mx100 ip ef=128 Q=10000" timeout 600 $NCU -k regex:fnb_search_kernel -c 1 -o /tmp/p_cfg2 python tools/ncu_one.py cfg2 --ef 128 > gpurun_out/r2ncu_cfg2.log 2>&1
python tools/ncu_summary.py r2_cfg2 - /tmp/p_cfg2.ncu-rep --shape cfg2
NCU_WORKLOAD="cfg4s 400kx960 K=100 ef=300 Q=5000" timeout 600 $NCU -k regex:fnb_search_kernel -c 1 -o /tmp/p_cfg4 python tools/ncu_one.py cfg4s --ef 300 > gpurun_out/r2ncu_cfg4.log 2>&1
python tools/ncu_summary.py r2_cfg4s - /tmp/p_cfg4.ncu-rep --shape cfg4s
# 6. the latency kernel: one search_single launch
NCU_WORKLOAD="cfg1 ef=100 search_single" timeout 600 $NCU -k regex:fnb_search_cta_kernel -c 1 -o /tmp/p_lat python tools/ncu_one.py cfg1 --single --launches 4 > gpurun_out/r2ncu_lat.log 2>&1
python tools/ncu_summary.py r2_lat - /tmp/p_lat.ncu-rep --name cta_kernel --shape search_single --by-samples
# 7. construction kernels (a mid-build batch of a 1M x 128 build) and the exchange + merge kernel (one rank)
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:build_select_kernel -s 100 -c 1 -o /tmp/p_bsel python tools/ncu_build.py > gpurun_out/r2ncu_bsel.log 2>&1
python tools/ncu_summary.py r2_build_select - /tmp/p_bsel.ncu-rep --name build_select_kernel --no-traffic
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:build_prune_kernel -s 100 -c 1 -o /tmp/p_bprune python tools/ncu_build.py > gpurun_out/r2ncu_bprune.log 2>&1
python tools/ncu_summary.py r2_build_prune - /tmp/p_bprune.ncu-rep --name build_prune_kernel --no-traffic
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:exchange_merge_kernel -s 2 -c 1 -o /tmp/p_ex python tools/sanitizer_cases.py --child exchange > gpurun_out/r2ncu_ex.log 2>&1
python tools/ncu_summary.py r2_exchange - /tmp/p_ex.ncu-rep --name exchange_merge_kernel --no-traffic
# 8. the launches of smoke() (what the driver's GPUTEST lists)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file /tmp/smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ncu_smoke.log 2>&1
python tools/ncu_summary.py r2_smoke /tmp/smoke.csv -
cp /tmp/p_cfg1.ncu-rep gpurun_out/r2_cfg1_search_kernel.ncu-rep
ls -la gpurun_out/profiles_r2
