#!/bin/bash
# Round-2 session b: GPU tests with the fused filter/CSR path, A/B of the gather walk, bench line, w100 ncu capture.
O=gpurun_out/r2b; mkdir -p $O
timeout -s KILL 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/pytest.log
QB_WINDOWS=100,250 timeout 900 python tools/ab.py run 100000 > $O/ab_walk.log 2>&1
timeout 1500 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"smooth_kernel|center_rows|colsum_dense|filter_|indptr_scan|gene_values" -o $O/step_w100 python tools/one_step.py 100000 100 > $O/ncu_w100.log 2>&1
tail -n 6 $O/pytest.log; cat $O/ab_walk.log; tail -c 600 $O/bench_n1.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2b/bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print("e2e", d["e2e"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "ms_per_launch")})
for k, v in d["sub"].items():
    print(k, v["value"], v["ms_per_step"], v["roofline"]["frac"], v["roofline"]["ms_per_launch"])
print("cpu", d["cpu_baseline"]["value"])
PY
