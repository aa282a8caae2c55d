#!/bin/bash
# w250 row pairs (global C scratch) — parity then timing
mkdir -p gpurun_out/s2
timeout 900 python -m pytest tests/test_infercnv_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/s2/pytest.log
ICNV_SMOOTH_ROWS=1 timeout 600 python -m pytest tests/test_infercnv_gpu.py -m gpu -q -x -k "250 or golden" 2>&1 | tail -8 > gpurun_out/s2/pytest_rows1.log
QB_WINDOWS=250,100 timeout 300 python tools/quick_bench.py 100000 > gpurun_out/s2/qb.log 2>&1
ICNV_SMOOTH_ROWS=1 QB_WINDOWS=250 timeout 300 python tools/quick_bench.py 100000 > gpurun_out/s2/qb_rows1.log 2>&1
tail -n 6 gpurun_out/s2/*.log
