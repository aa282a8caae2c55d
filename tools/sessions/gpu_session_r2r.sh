#!/bin/bash
O=gpurun_out/r2r; mkdir -p $O
timeout 600 python tools/leiden_profile.py 1000000 > $O/leiden_profile.log 2>&1
tail -4 $O/leiden_profile.log
