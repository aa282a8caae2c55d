#!/bin/bash
O=gpurun_out/r2s; mkdir -p $O

timeout 600 python tools/leiden_profile.py 1000000 > $O/leiden_profile.log 2>&1
tail -n 3 $O/pytest.log; tail -16 $O/leiden_profile.log
