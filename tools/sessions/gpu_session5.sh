#!/bin/bash
mkdir -p gpurun_out/s5
timeout 600 python -m pytest tests/test_infercnv_gpu.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/s5/pytest.log
ICNV_SMOOTH_ROWS=1 timeout 600 python -m pytest tests/test_infercnv_gpu.py -m gpu -q -x -k "golden or bench_chunk or float64" 2>&1 | tail -3 > gpurun_out/s5/pytest_rows1.log
QB_WINDOWS=250,100 timeout 300 python tools/quick_bench.py 100000 > gpurun_out/s5/qb_default.log 2>&1
ICNV_SMOOTH_ROWS=1 QB_WINDOWS=100 timeout 300 python tools/quick_bench.py 100000 > gpurun_out/s5/qb_rows1_dbuf.log 2>&1
ICNV_SMOOTH_DBUF=0 ICNV_SMOOTH_ROWS=1 QB_WINDOWS=250,100 timeout 300 python tools/quick_bench.py 100000 > gpurun_out/s5/qb_rows1_nodbuf.log 2>&1
timeout 300 python tools/timeline.py 29600 250 > gpurun_out/s5/timeline_w250_dbuf.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s5/bench.json 2> gpurun_out/s5/bench.err
tail -n 5 gpurun_out/s5/*.log; tail -c 1500 gpurun_out/s5/bench.err; head -c 3000 gpurun_out/s5/bench.json
