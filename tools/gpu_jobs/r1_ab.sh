mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/ab_gpu.txt
timeout 900 python tools/ab_probe.py --libs "${LIBS:-base=flatnav_b200/libflatnav_b200.so,spec=variants/libspec.so,specadj=variants/libspecadj.so}" --cases "${CASES:-cfg1,cfg2,u8,cfg4s}" --out gpurun_out/ab.json 2>&1 | tee gpurun_out/ab.log | tail -30
