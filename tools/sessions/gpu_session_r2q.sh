#!/bin/bash
# Round-2 session q: compute-sanitizer (memcheck + racecheck) over every kernel family on small inputs.
O=gpurun_out/r2q; mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/memcheck.log 2>&1
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_small.py > $O/racecheck.log 2>&1
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" $O/memcheck.log $O/racecheck.log | head -20; tail -3 $O/memcheck.log
