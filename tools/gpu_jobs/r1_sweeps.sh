# refresh the ef sweeps of the BASELINE configs with the current kernel (profiles/sweeps/)
mkdir -p gpurun_out
for spec in "cfg2 reference" "cfg1 reference" "cfg5shard gpu" "cfg4 gpu"; do
  set -- $spec
  ( time timeout 400 python tools/sweep.py $1 --builder $2 --out gpurun_out/sweep_$1.json ) > gpurun_out/sweep_$1.log 2>&1
  grep "\[sweep\] {" gpurun_out/sweep_$1.log | cut -c1-200; tail -3 gpurun_out/sweep_$1.log | grep real
done
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log
