#!/bin/bash
# Round-2 final evidence (one GPU): GPU suite, smoke, the driver's bench line, reference arm, launch list.
O=gpurun_out/r2y; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/box.txt
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 1500 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
tail -n 3 $O/pytest.log; tail -n 2 $O/smoke.log; tail -c 400 $O/bench_n1.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2y/bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print("e2e", {k: d["e2e"][k] for k in ("value", "seconds_per_step", "h2d_only_seconds", "h2d_bytes_per_step", "d2h_bytes_per_step")})
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "ms_per_launch")})
for k, v in d["sub"].items():
    if "roofline" in v: print(k, v["value"], v["ms_per_step"], v["roofline"]["frac"], v["roofline"]["ms_per_launch"])
    else: print(k, v)
print("cpu", d["cpu_baseline"]["value"])
PY
