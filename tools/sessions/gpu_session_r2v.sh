#!/bin/bash
O=gpurun_out/r2v; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -k "workflow_sequence" 2>&1 | tail -4 > $O/pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"community_sweep_kernel|community_totals_kernel" -s 40 -c 4 -o $O/sweep python tools/leiden_profile.py 1000000 > $O/ncu.log 2>&1
tail -n 3 $O/pytest.log; tail -3 $O/ncu.log
