#!/bin/bash
# Round-2 session i: new tests (edge cases, ITH selection path), graph workload (configs[4]) at 1M cells on one GPU.
O=gpurun_out/r2i; mkdir -p $O
timeout -s KILL 900 python -m pytest tests/test_infercnv_gpu.py -m gpu -q -x -k "edge or ith or loader or cnv_score" 2>&1 | tail -6 > $O/pytest.log
timeout 1500 python bench.py --workloads dense100,graph --cells-total 1000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/bench_graph1m.json 2> $O/bench_graph1m.err
tail -n 4 $O/pytest.log; tail -c 300 $O/bench_graph1m.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2i/bench_graph1m.json"))
print(json.dumps(d["sub"]["graph"], indent=1))
PY
