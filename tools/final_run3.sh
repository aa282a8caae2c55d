set -x
timeout -s KILL 300 python -u -m pytest tests -m gpu -x -q --timeout 150 --timeout-method=thread 2>&1 | tail -3
python bench.py > gpurun_out/bench_dense100.json 2> gpurun_out/bench_dense100.err
QB_GENEVALS=1 QB_WINDOWS=100 QB_REPS=5 python tools/quick_bench.py 100000 2>&1 | grep -E "gene_values|smooth_ms" | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
tail -c 300 gpurun_out/bench_dense100.json
