# round 2, job 2: source-level ncu of the round-1 library and the new-merge build (per-line CSVs only: .ncu-rep files are
# ~20 MB each and gpurun_out is capped at 64 MiB); A/B of the balanced grid
mkdir -p gpurun_out
timeout 600 python tools/ab_probe.py --libs "prev=variants/libprev.so,new_nobal=flatnav_b200/libflatnav_b200.so@FNB_BALANCED_GRID=0,new=flatnav_b200/libflatnav_b200.so" --cases "cfg1,u8" --out gpurun_out/r2j2_ab.json 2>&1 | tee gpurun_out/r2j2_ab.log | tail -12
for v in prev new; do
  lib=flatnav_b200/libflatnav_b200.so; [ $v = prev ] && lib=variants/libprev.so
  for c in u8 cfg1; do
    FNB_LIB_PATH=$PWD/$lib timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:fnb_search_kernel -c 1 -f -o /tmp/r2j2_${v}_${c} python tools/ncu_one.py $c > gpurun_out/r2j2_ncu_${v}_${c}.log 2>&1
    ncu -i /tmp/r2j2_${v}_${c}.ncu-rep --page source --print-source cuda --csv > gpurun_out/r2j2_${v}_${c}_source.csv 2>/dev/null
    ncu -i /tmp/r2j2_${v}_${c}.ncu-rep --page raw --csv > gpurun_out/r2j2_${v}_${c}_raw.csv 2>/dev/null
  done
done
cp /tmp/r2j2_new_u8.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out | grep r2j2
