#!/usr/bin/env bash
# Everything under profiles/ comes from commands like these, run on a B200 box through `gpurun -- bash tools/round_evidence.sh`
# (outputs land in gpurun_out/; the summaries are then made here with tools/launch_list.py and tools/ncu_summary.py).
# The exact per-session scripts of round 2 are kept under tools/sessions/.
set -x
O=gpurun_out/evidence; mkdir -p $O
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err                       # dense100 + dense250 + csr100 + graph + e2e + cpu baseline
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
# launch list of a short bench run (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv \
    python bench.py --cells-total 100000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --workloads dense100,dense250,csr100 > /dev/null 2>&1
# one launch of every kernel at the bench size (kernel names carry no namespace in ncu's -k filter)
ncu --set full --clock-control none --import-source on \
    -k regex:"smooth_kernel|center_rows|colsum_dense|filter_|indptr_scan|gene_values" -o $O/step_w100 python tools/one_step.py 100000 100 > $O/one_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"smooth_kernel" -c 1 -o $O/step_w250 python tools/one_step.py 100000 250 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"smooth_csr|colsum_csr" -c 8 -o $O/step_csr python tools/csr_one.py 100000 100 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn -c 3 -o $O/knn python tools/knn_one.py 65536 50 > /dev/null 2>&1
# sanitizers on small inputs
compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > $O/memcheck.log 2>&1
compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_small.py > $O/racecheck.log 2>&1
tail -c 300 $O/bench_n1.json
