set -x
python bench.py > gpurun_out/bench_dense100.json 2> gpurun_out/bench_dense100.err
ICNV_SPLIT_ROWS=0 python bench.py --steps 100 --no-e2e --no-cpu-baseline > gpurun_out/bench_split0.json 2>/dev/null
python bench.py --workload dense250 --steps 50 --no-cpu-baseline > gpurun_out/bench_dense250.json 2>/dev/null
python bench.py --workload csr100 --steps 50 --no-cpu-baseline > gpurun_out/bench_csr100.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:icnv -o gpurun_out/step_r1b python tools/one_step.py 100000 100 > gpurun_out/one_step.log 2>&1
ncu --set full --clock-control none -k regex:smooth_kernel -o gpurun_out/smooth_w250 python tools/one_step.py 50000 250 > /dev/null 2>&1
tail -c 300 gpurun_out/bench_dense100.json; tail -c 200 gpurun_out/bench_split0.json
