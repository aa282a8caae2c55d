#!/bin/bash
mkdir -p gpurun_out/s7
timeout -s KILL 240 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -k "knn_is_exact" 2>&1 | tail -25 > gpurun_out/s7/pytest_knn_small.log
timeout -s KILL 240 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -k "knn_tensor_core" 2>&1 | tail -25 > gpurun_out/s7/pytest_knn_scale.log
timeout -s KILL 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/s7/pytest_all.log
timeout 600 python bench.py --workloads csr100 --cells-total 200000 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s7/bench_csr_sparse.json 2> gpurun_out/s7/bench_csr_sparse.err
tail -n 8 gpurun_out/s7/pytest_knn_small.log gpurun_out/s7/pytest_knn_scale.log; tail -n 4 gpurun_out/s7/pytest_all.log
