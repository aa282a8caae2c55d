#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
QB_WINDOWS=250 timeout 600 python tools/ab.py run 100000 > $O/ab_threads.log 2>&1
cat $O/ab_threads.log
