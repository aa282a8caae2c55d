#!/bin/bash
# round-2 session 1: full GPU suite with the tutorial-shaped goldens enabled, then the banded kernel (first run ever)
mkdir -p gpurun_out/s1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1/smi.txt
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -80 > gpurun_out/s1/pytest_default.log
ICNV_SMOOTH_BANDS=2 timeout 600 python -m pytest tests/test_infercnv_gpu.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/s1/pytest_banded.log
timeout 300 python tools/quick_bench.py 100000 > gpurun_out/s1/qb_default.log 2>&1
ICNV_SMOOTH_BANDS=2 QB_WINDOWS=100 timeout 300 python tools/quick_bench.py 100000 > gpurun_out/s1/qb_banded.log 2>&1
tail -5 gpurun_out/s1/*.log
