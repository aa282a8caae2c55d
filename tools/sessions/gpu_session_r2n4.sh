#!/bin/bash
N=${1:-4}
O=gpurun_out/r2n$N; mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
echo "rc=$?"; head -c 200 $O/bench.json; python - <<PY
import json
d = json.loads([l for l in open("$O/bench.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "frac", d["roofline"]["frac"])
print("e2e", {k: d["e2e"][k] for k in ("value", "seconds_per_step", "h2d_only_seconds", "h2d_only_GBps_per_rank")})
for k, v in d["sub"].items():
    print(k, v["value"], v.get("ms_per_step"), v.get("stage_seconds"))
PY
