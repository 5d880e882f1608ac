bash tools/gpu_jobs/r2_final.sh
bash tools/gpu_jobs/r2_sanitizer.sh
