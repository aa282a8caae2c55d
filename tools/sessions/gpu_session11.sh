#!/bin/bash
mkdir -p gpurun_out/s11
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/s11/pytest.log
timeout 300 python tools/knn_one.py 65536 50 > gpurun_out/s11/knn_one.log 2>&1
timeout 300 python tools/knn_one.py 262144 50 >> gpurun_out/s11/knn_one.log 2>&1
timeout 600 python tools/knn_one.py 1000000 50 >> gpurun_out/s11/knn_one.log 2>&1
QB_GENEVALS=1 QB_WINDOWS=100 timeout 300 python tools/quick_bench.py 50000 > gpurun_out/s11/qb_genevals.log 2>&1
timeout 1200 python bench.py --workloads dense100,graph --cells-total 1000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/s11/bench_graph1m.json 2> gpurun_out/s11/bench_graph1m.err
tail -n 4 gpurun_out/s11/pytest.log; cat gpurun_out/s11/knn_one.log; grep gene_values gpurun_out/s11/qb_genevals.log; tail -c 400 gpurun_out/s11/bench_graph1m.err
