#!/bin/bash
mkdir -p gpurun_out/s3
QB_WINDOWS=250 timeout 300 python tools/quick_bench.py 100000 > gpurun_out/s3/qb_pairs.log 2>&1
ICNV_SMOOTH_ROWS=1 ICNV_LIB_PATH=infercnvpy_b200/ab/libicnv_cg1.so QB_WINDOWS=250 timeout 300 python tools/quick_bench.py 100000 > gpurun_out/s3/qb_cg1.log 2>&1
ICNV_SMOOTH_ROWS=1 ICNV_LIB_PATH=infercnvpy_b200/ab/libicnv_cg1.so timeout 300 python -m pytest tests/test_infercnv_gpu.py -m gpu -q -x -k "250 or golden" 2>&1 | tail -3 > gpurun_out/s3/pytest_cg1.log
timeout 300 python tools/timeline.py 29600 250 > gpurun_out/s3/timeline_pairs.log 2>&1
ICNV_SMOOTH_ROWS=1 timeout 300 python tools/timeline.py 29600 250 > gpurun_out/s3/timeline_rows1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smooth_kernel -c 1 -o gpurun_out/s3/w250_pairs python tools/one_step.py 29600 250 > gpurun_out/s3/ncu.log 2>&1
tail -n 4 gpurun_out/s3/*.log
