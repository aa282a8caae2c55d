#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
timeout -s KILL 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -k "knn or workflow" 2>&1 | tail -4 > $O/pytest.log
timeout 300 python tools/knn_one.py 65536 50 > $O/knn_one.log 2>&1
timeout 300 python tools/knn_one.py 262144 50 >> $O/knn_one.log 2>&1
timeout 600 python tools/knn_one.py 1000000 50 >> $O/knn_one.log 2>&1
tail -n 3 $O/pytest.log; grep -v Warn $O/knn_one.log
