#!/bin/bash
O=gpurun_out/r2k; mkdir -p $O
echo "== default (row pairs, permuted walk)" > $O/ab_single.log
QB_WINDOWS=100 QB_REPS=10 timeout 300 python tools/quick_bench.py 100000 2>&1 | grep '"window"' | python -c "import sys,json; [print(json.loads(l)['launch'], json.loads(l)['smooth_ms'], json.loads(l)['smooth_frac']) for l in sys.stdin]" >> $O/ab_single.log
echo "== ICNV_SMOOTH_ROWS=1 (single row, double-buffered partials, permuted walk)" >> $O/ab_single.log
ICNV_SMOOTH_ROWS=1 QB_WINDOWS=100 QB_REPS=10 timeout 300 python tools/quick_bench.py 100000 2>&1 | grep '"window"' | python -c "import sys,json; [print(json.loads(l)['launch'], json.loads(l)['smooth_ms'], json.loads(l)['smooth_frac']) for l in sys.stdin]" >> $O/ab_single.log
echo "== ICNV_SMOOTH_ROWS=1 + natural walk everywhere" >> $O/ab_single.log
ICNV_LIB_PATH=infercnvpy_b200/ab/libicnv_nat2.so ICNV_SMOOTH_ROWS=1 QB_WINDOWS=100 QB_REPS=10 timeout 300 python tools/quick_bench.py 100000 2>&1 | grep '"window"' | python -c "import sys,json; [print(json.loads(l)['launch'], json.loads(l)['smooth_ms'], json.loads(l)['smooth_frac']) for l in sys.stdin]" >> $O/ab_single.log
cat $O/ab_single.log
