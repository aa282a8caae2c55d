#!/bin/bash
# Round-2 session e: full GPU suite (loader, umap / tsne, delta kernel with warp-uniform skip), CSR timings.
O=gpurun_out/r2e; mkdir -p $O
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v Warning | tail -25 > $O/pytest.log
timeout 300 python tools/csr_one.py 100000 100 > $O/csr_one.log 2>&1
timeout 300 python tools/csr_one.py 100000 250 >> $O/csr_one.log 2>&1
ICNV_CSR_DELTA=0 timeout 300 python tools/csr_one.py 100000 250 >> $O/csr_one.log 2>&1
tail -n 14 $O/pytest.log; grep -v Warn $O/csr_one.log
