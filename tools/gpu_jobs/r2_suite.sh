# round 2: the GPU suite on the committed tree (default configuration)
mkdir -p gpurun_out
( time timeout 110 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu.log | head -3
