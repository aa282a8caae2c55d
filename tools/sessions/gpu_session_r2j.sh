#!/bin/bash
O=gpurun_out/r2j; mkdir -p $O
timeout 600 python tools/e2e_tail.py 250000 > $O/e2e_tail.log 2>&1
grep -v "Using mean" $O/e2e_tail.log | tail -8
