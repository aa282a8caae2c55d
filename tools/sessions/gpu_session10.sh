#!/bin/bash
mkdir -p gpurun_out/s10
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/s10/pytest.log
timeout 300 python tools/knn_one.py 65536 50 > gpurun_out/s10/knn_one.log 2>&1
timeout 300 python tools/knn_one.py 262144 50 >> gpurun_out/s10/knn_one.log 2>&1
timeout 600 python tools/knn_one.py 1000000 50 >> gpurun_out/s10/knn_one.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_mma_kernel -c 1 -o gpurun_out/s10/knn python tools/knn_one.py 65536 50 > gpurun_out/s10/ncu_knn.log 2>&1
tail -n 4 gpurun_out/s10/pytest.log; cat gpurun_out/s10/knn_one.log
