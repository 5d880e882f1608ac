# round 2, job 1: new register-scatter merge vs the round-1 library (variants/libprev.so), and source-level ncu of both
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2j1_gpu.txt; nproc >> gpurun_out/r2j1_gpu.txt
timeout 900 python tools/ab_probe.py --libs "prev=variants/libprev.so,new=flatnav_b200/libflatnav_b200.so" --cases "cfg1,u8,cfg1big,u8big,cfg2" --out gpurun_out/r2j1_ab.json 2>&1 | tee gpurun_out/r2j1_ab.log | tail -30
for v in prev new; do
  lib=flatnav_b200/libflatnav_b200.so; [ $v = prev ] && lib=variants/libprev.so
  for c in u8 cfg1; do
    FNB_LIB_PATH=$PWD/$lib timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:fnb_search_kernel -c 1 -f -o gpurun_out/r2j1_${v}_${c} python tools/ncu_one.py $c > gpurun_out/r2j1_ncu_${v}_${c}.log 2>&1
    tail -2 gpurun_out/r2j1_ncu_${v}_${c}.log
  done
done
ls -la gpurun_out | grep r2j1
