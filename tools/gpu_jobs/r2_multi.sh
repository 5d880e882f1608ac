# round 2, N GPUs (gpurun --gpus N): the world-size tests, both bench arms under torchrun exactly as the driver launches them
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2m${N}_gpus.txt; nproc >> gpurun_out/r2m${N}_gpus.txt; free -g | head -2 >> gpurun_out/r2m${N}_gpus.txt
if [ -z "$SKIP_TESTS" ]; then ( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/r2m${N}_pytest.log 2>&1; tail -5 gpurun_out/r2m${N}_pytest.log; fi
if [ -z "$SKIP_REF" ]; then timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus $N --steps 10 --warmup 3 > gpurun_out/r2m${N}_bench_ref.json 2> gpurun_out/r2m${N}_bench_ref.err; cat gpurun_out/r2m${N}_bench_ref.json | cut -c1-400; fi
( time NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2m${N}_bench.json 2> gpurun_out/r2m${N}_bench.err ); tail -c 2500 gpurun_out/r2m${N}_bench.json; grep "bench " gpurun_out/r2m${N}_bench.err | tail -5; grep -c "AllGather" gpurun_out/r2m${N}_bench.err
tail -c 20000 gpurun_out/r2m${N}_bench.err > gpurun_out/r2m${N}_bench.err.tail; rm gpurun_out/r2m${N}_bench.err
