#!/bin/bash
# Round-2 last confirmation: the driver's GPU commands on the final tree.
O=gpurun_out/r2x; mkdir -p $O
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > $O/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
tail -n 3 $O/pytest.log; tail -n 1 $O/smoke.log
