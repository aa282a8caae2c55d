#!/bin/bash
O=gpurun_out/r2w; mkdir -p $O
timeout -s KILL 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -s -k "leiden or community or workflow" 2>&1 | grep -v Warn | tail -8 > $O/pytest.log
timeout 600 python tools/leiden_profile.py 1000000 > $O/leiden_profile.log 2>&1
tail -n 6 $O/pytest.log; grep -A4 "move_tol=0.001" $O/leiden_profile.log; grep "move_tol=0:" $O/leiden_profile.log
