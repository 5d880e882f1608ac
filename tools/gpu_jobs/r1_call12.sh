# occupancy A/B at large batches (no wave-quantisation tail)
mkdir -p gpurun_out
( time timeout 600 python tools/ab_probe.py --reps 5 \
   --libs "base=flatnav_b200/libflatnav_b200.so,short5=variants/libshort5.so,short7=variants/libshort7.so,short8=variants/libshort8.so" \
   --cases "${CASES:-cfg1big,cfg2big,u8big}" --out gpurun_out/ab_big.json ) > gpurun_out/ab_big.log 2>&1
tail -25 gpurun_out/ab_big.log
