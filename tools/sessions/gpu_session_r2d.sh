#!/bin/bash
# Round-2 session d: delta CSR kernel with SoA limb arrays.
O=gpurun_out/r2d; mkdir -p $O
timeout -s KILL 900 python -m pytest tests/test_infercnv_gpu.py -m gpu -q -x -s -k "csr or scale or bench_chunk or multi_block or golden or known" 2>&1 | grep -v Warning | tail -12 > $O/pytest.log
timeout 300 python tools/csr_one.py 100000 100 > $O/csr_one.log 2>&1
timeout 300 python tools/csr_one.py 100000 250 >> $O/csr_one.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"smooth_csr_delta" -c 1 -o $O/step_csr python tools/csr_one.py 100000 100 > $O/ncu_csr.log 2>&1
timeout 600 python tools/e2e_breakdown.py > $O/e2e_breakdown.log 2>&1
tail -n 6 $O/pytest.log; grep -v Warn $O/csr_one.log; tail -30 $O/e2e_breakdown.log
