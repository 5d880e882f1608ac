# bench.py under torchrun must put exactly one line (the JSON result) on stdout
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/stdout_n2.txt 2> gpurun_out/stderr_n2.txt
wc -l gpurun_out/stdout_n2.txt; head -c 200 gpurun_out/stdout_n2.txt; echo; grep -c "NCCL version" gpurun_out/stderr_n2.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/stdout_ref_n2.txt 2> gpurun_out/stderr_ref_n2.txt
wc -l gpurun_out/stdout_ref_n2.txt; head -c 200 gpurun_out/stdout_ref_n2.txt; echo
