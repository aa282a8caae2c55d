#!/bin/bash
mkdir -p gpurun_out/s8
timeout -s KILL 600 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -s 2>&1 | tail -30 > gpurun_out/s8/pytest_graph.log
cat > /tmp/csr_one.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch, infercnvpy_b200 as cnv
from infercnvpy_b200._engine import DevicePlan
from infercnvpy_b200._layout import build_layout
N=int(sys.argv[1]); dev=torch.device("cuda",0)
var=cnv.datasets.synthetic_var(20000,seed=0); Xd=cnv.datasets.device_counts(N,20000,dev,seed=1000)
csr=Xd.to_sparse_csr(); t=(csr.crow_indices().to(torch.int64),csr.col_indices().to(torch.int32),csr.values())
with DevicePlan(build_layout(var,100,10),dev) as plan:
    s,c=plan.colsum(t); plan.set_reference(plan.mean_from_sums(s,c))
    for _ in range(3): tmp=plan.smooth(t,3.0)
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); tmp=plan.smooth(t,3.0); b.record(); torch.cuda.synchronize(); print("smooth_csr ms", a.elapsed_time(b), "rows", N)
    a.record(); s,c=plan.colsum(t); b.record(); torch.cuda.synchronize(); print("colsum_csr ms", a.elapsed_time(b))
PY
timeout 300 python /tmp/csr_one.py 100000 > gpurun_out/s8/csr_one.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smooth_csr_kernel -c 1 -o gpurun_out/s8/csr_sparse python /tmp/csr_one.py 29600 > gpurun_out/s8/ncu.log 2>&1
tail -n 12 gpurun_out/s8/pytest_graph.log; cat gpurun_out/s8/csr_one.log
