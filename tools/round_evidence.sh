#!/usr/bin/env bash
# Everything under profiles/ comes from these commands, run on a B200 box through `gpurun -- bash tools/round_evidence.sh`
# (outputs land in gpurun_out/; the summaries are then made here with tools/launch_list.py and tools/ncu_summary.py).
set -x
timeout -s KILL 300 python -u -m pytest tests -m gpu -x -q --timeout 150 --timeout-method=thread 2>&1 | tail -3
python bench.py > gpurun_out/bench_dense100.json 2> gpurun_out/bench_dense100.err
python bench.py --workload dense250 --steps 50 --no-cpu-baseline > gpurun_out/bench_dense250.json 2>/dev/null
python bench.py --workload csr100 --steps 50 --no-cpu-baseline > gpurun_out/bench_csr100.json 2>/dev/null
# launch list of a short bench run (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
# one launch of every kernel at the bench size
ncu --set full --clock-control none --import-source on \
    -k regex:"smooth_kernel|center_rows_kernel|colsum_dense_kernel|apply_threshold_kernel|gene_values_kernel|dense_to_csr_kernel" \
    -o gpurun_out/step python tools/one_step.py 100000 100 > gpurun_out/one_step.log 2>&1
# sanitizers on small inputs
compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/memcheck.log 2>&1
compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/racecheck.log 2>&1
tail -c 300 gpurun_out/bench_dense100.json
