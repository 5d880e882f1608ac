# ncu --set full of the latency variant (one query, one warp on an idle GPU)
mkdir -p gpurun_out
cat > /tmp/lat_one.py <<'P'
import sys, numpy as np
sys.path.insert(0, ".")
import flatnav_b200
from flatnav_b200 import synthetic
from tools.workload import ensure_index
path, _ = ensure_index("latent", 1_000_000, 128, "l2", 32, 100, builder="gpu")
ix = flatnav_b200.index.IndexL2Float.load_index(path, devices=[0])
q = synthetic.make("latent", 64, 128, queries=True)
import torch
for i in range(8):
    ix.search_single(q[i], 10, 100)
torch.cuda.synchronize()
torch.cuda.profiler.start()   # construction above launches the same kernel template: profile the queries only
for i in range(8, 40):
    ix.search_single(q[i], 10, 100)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
P
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fnb_search_kernel -s 8 -c 1 -f -o gpurun_out/prof_lat python /tmp/lat_one.py > gpurun_out/ncu_lat.log 2>&1
tail -3 gpurun_out/ncu_lat.log; ls -la gpurun_out/prof_lat.ncu-rep
