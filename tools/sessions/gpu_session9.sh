#!/bin/bash
mkdir -p gpurun_out/s9
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/s9/pytest.log
cat > /tmp/csr_one.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch, infercnvpy_b200 as cnv
from infercnvpy_b200._engine import DevicePlan
from infercnvpy_b200._layout import build_layout
N=int(sys.argv[1]); w=int(sys.argv[2]); dev=torch.device("cuda",0)
var=cnv.datasets.synthetic_var(20000,seed=0); Xd=cnv.datasets.device_counts(N,20000,dev,seed=1000)
csr=Xd.to_sparse_csr(); t=(csr.crow_indices().to(torch.int64),csr.col_indices().to(torch.int32),csr.values())
with DevicePlan(build_layout(var,w,10),dev) as plan:
    s,c=plan.colsum(t); plan.set_reference(plan.mean_from_sums(s,c))
    for _ in range(3): tmp=plan.smooth(t,3.0)
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); tmp=plan.smooth(t,3.0); b.record(); torch.cuda.synchronize(); print("window", w, "smooth_csr ms", a.elapsed_time(b), "rows", N)
    a.record(); s,c=plan.colsum(t); b.record(); torch.cuda.synchronize(); print("colsum_csr ms", a.elapsed_time(b))
PY
timeout 300 python /tmp/csr_one.py 100000 100 > gpurun_out/s9/csr_one.log 2>&1
timeout 300 python /tmp/csr_one.py 100000 250 >> gpurun_out/s9/csr_one.log 2>&1
timeout 300 python tools/knn_one.py 65536 50 > gpurun_out/s9/knn_one.log 2>&1
timeout 300 python tools/knn_one.py 262144 50 >> gpurun_out/s9/knn_one.log 2>&1
timeout 900 python bench.py --workloads dense100,graph --cells-total 200000 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/s9/bench_graph200k.json 2> gpurun_out/s9/bench_graph200k.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_mma_kernel -c 1 -o gpurun_out/s9/knn python tools/knn_one.py 65536 50 > gpurun_out/s9/ncu_knn.log 2>&1
tail -n 5 gpurun_out/s9/pytest.log; cat gpurun_out/s9/csr_one.log gpurun_out/s9/knn_one.log | grep -v Warn; tail -c 600 gpurun_out/s9/bench_graph200k.err
