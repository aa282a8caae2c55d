#!/bin/bash
# Round-2 session c: delta CSR kernel + natural walk default: tests, timings, bench line, ncu captures.
O=gpurun_out/r2c; mkdir -p $O
timeout -s KILL 1200 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v Warning | tail -25 > $O/pytest.log
timeout 300 python tools/csr_one.py 100000 100 > $O/csr_one.log 2>&1
timeout 300 python tools/csr_one.py 100000 250 >> $O/csr_one.log 2>&1
ICNV_CSR_DELTA=0 timeout 300 python tools/csr_one.py 100000 100 >> $O/csr_one.log 2>&1
timeout 1500 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"smooth_csr" -c 6 -o $O/step_csr python tools/csr_one.py 100000 100 > $O/ncu_csr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"smooth_kernel" -c 1 -o $O/step_w250 python tools/one_step.py 100000 250 > $O/ncu_w250.log 2>&1
tail -n 8 $O/pytest.log; grep -v Warn $O/csr_one.log; tail -c 600 $O/bench_n1.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c/bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print("e2e", d["e2e"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "ms_per_launch")})
for k, v in d["sub"].items():
    print(k, v["value"], v["ms_per_step"], v["roofline"]["frac"], v["roofline"]["ms_per_launch"])
print("cpu", d["cpu_baseline"]["value"])
PY
