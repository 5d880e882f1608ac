# multi-GPU: world-size tests, bench.py under torchrun, dataset-sharded bench.  N = number of GPUs of the call.
N=${N:-2}
NSHARD=${NSHARD:-4000000}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/multi_gpus.txt; nproc >> gpurun_out/multi_gpus.txt
[ -n "$SKIP_TESTS" ] || ( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/pytest_multi_n$N.log 2>&1; tail -4 gpurun_out/pytest_multi_n$N.log
[ -n "$SKIP_REF" ] || python bench.py --impl reference --steps 5 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; cut -c1-200 gpurun_out/bench_ref_n$N.json
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 3 ) > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/bench_sharded.py --n-shard $NSHARD --steps 50 ) > gpurun_out/sharded_n$N.jsonl 2> gpurun_out/sharded_n$N.err; cut -c1-400 gpurun_out/sharded_n$N.jsonl; tail -3 gpurun_out/sharded_n$N.err
