#!/bin/bash
# Round-2 session p (8 GPUs): the driver's scaling command at N=8.
O=gpurun_out/r2p; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err
echo "rc=$?"; grep -v "Using mean\|^$\|OMP_NUM\|\*\*\*\*" $O/bench_n8.err | tail -15; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2p/bench_n8.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "scaling", "clocks")})
print("e2e", {k: d["e2e"][k] for k in ("value", "seconds_per_step", "h2d_only_seconds", "h2d_only_GBps_per_rank", "cells_per_step_all_ranks")})
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "ms_per_launch")})
for k, v in d["sub"].items():
    if "roofline" in v: print(k, v["value"], v["ms_per_step"], v["roofline"]["frac"])
    else: print(k, {a: v[a] for a in v if a != "config"})
PY
head -20 $O/topo.txt
