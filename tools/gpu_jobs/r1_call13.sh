# per-launch occupancy plan (24 vs 28 warps per SM): A/B of the automatic choice against both forced plans, parity under the forced dense plan
mkdir -p gpurun_out
L=flatnav_b200/libflatnav_b200.so
( time timeout 600 python tools/ab_probe.py --reps 5 --libs "auto=$L,sparse=$L@FNB_DENSE=0,dense=$L@FNB_DENSE=1" \
   --cases "${CASES:-cfg1,cfg1big,cfg2,cfg2big,u8,u8big,cfg3s}" --out gpurun_out/ab_dense.json ) > gpurun_out/ab_dense.log 2>&1
tail -30 gpurun_out/ab_dense.log
( time FNB_DENSE=1 FNB_LAT=0 timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_dense.log 2>&1; tail -4 gpurun_out/pytest_gpu_dense.log
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
