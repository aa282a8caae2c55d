#!/bin/bash
# Round-2 session f: split phase 3 of the double-buffered single-row kernels (window 250): parity + A/B.
O=gpurun_out/r2f; mkdir -p $O
timeout -s KILL 900 python -m pytest tests/test_infercnv_gpu.py -m gpu -q -x -k "golden or bench_chunk or scale or float64 or general or known" 2>&1 | tail -6 > $O/pytest.log
QB_WINDOWS=250,100 timeout 600 python tools/ab.py run 100000 > $O/ab_p3split.log 2>&1
tail -n 4 $O/pytest.log; cat $O/ab_p3split.log
