#!/bin/bash
O=gpurun_out/r2u; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -k "gene_values or pca or multi_block or general" 2>&1 | tail -5 > $O/pytest.log
QB_GENEVALS=1 QB_WINDOWS=100,250 timeout 300 python tools/quick_bench.py 50000 > $O/qb_genevals.log 2>&1
tail -n 4 $O/pytest.log; grep gene_values $O/qb_genevals.log
