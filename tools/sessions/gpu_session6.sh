#!/bin/bash
mkdir -p gpurun_out/s6
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/s6/pytest.log
QB_WINDOWS=250,100 timeout 300 python tools/quick_bench.py 100000 > gpurun_out/s6/qb_t768.log 2>&1
ICNV_SMOOTH_ROWS=1 QB_WINDOWS=100 timeout 300 python tools/quick_bench.py 100000 > gpurun_out/s6/qb_t768_rows1.log 2>&1
for v in t512 t1024; do
ICNV_LIB_PATH=infercnvpy_b200/ab/libicnv_$v.so QB_WINDOWS=250 timeout 300 python tools/quick_bench.py 100000 > gpurun_out/s6/qb_$v.log 2>&1
ICNV_LIB_PATH=infercnvpy_b200/ab/libicnv_$v.so ICNV_SMOOTH_ROWS=1 QB_WINDOWS=100 timeout 300 python tools/quick_bench.py 100000 > gpurun_out/s6/qb_${v}_rows1.log 2>&1
done
ICNV_LIB_PATH=infercnvpy_b200/ab/libicnv_t1024.so timeout 300 python -m pytest tests/test_infercnv_gpu.py -m gpu -q -x -k "golden or bench_chunk" 2>&1 | tail -3 > gpurun_out/s6/pytest_t1024.log
timeout 600 python bench.py --workloads csr100 --cells-total 200000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s6/bench_csr_sparse.json 2> gpurun_out/s6/bench_csr_sparse.err
ICNV_CSR_SPARSE=0 timeout 600 python bench.py --workloads csr100 --cells-total 200000 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s6/bench_csr_densify.json 2> gpurun_out/s6/bench_csr_densify.err
tail -n 5 gpurun_out/s6/pytest*.log
