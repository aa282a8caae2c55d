#!/bin/bash
O=gpurun_out/r2aa; mkdir -p $O
timeout -s KILL 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x 2>&1 | tail -4 > $O/pytest.log
tail -n 3 $O/pytest.log
