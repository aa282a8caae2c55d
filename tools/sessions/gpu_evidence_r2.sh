#!/bin/bash
# Round-2 evidence run (one gpurun call): bench line at the metric's config, launch list, ncu --set full of the
# dominant kernels, then the GPU test suite.  Outputs under gpurun_out/r2a/.
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/box.txt
timeout 1500 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --cells-total 100000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:icnv -o $O/step_w100 python tools/one_step.py 100000 100 > $O/ncu_w100.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smooth -o $O/step_w250 python tools/one_step.py 100000 250 > $O/ncu_w250.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"smooth|colsum" -c 12 -o $O/step_csr python tools/csr_one.py 100000 100 > $O/ncu_csr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn -c 3 -o $O/knn python tools/knn_one.py 65536 50 > $O/ncu_knn.log 2>&1
timeout 300 python tools/quick_bench.py 100000 > $O/quick_bench.log 2>&1
timeout 300 python tools/csr_one.py 100000 100 > $O/csr_one.log 2>&1
timeout 300 python tools/csr_one.py 100000 250 >> $O/csr_one.log 2>&1
timeout 300 python tools/knn_one.py 262144 50 > $O/knn_one.log 2>&1
timeout -s KILL 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/pytest.log
ls -la $O; tail -n 5 $O/pytest.log; tail -c 1500 $O/bench_n1.err; head -c 3000 $O/bench_n1.json; cat $O/quick_bench.log $O/csr_one.log $O/knn_one.log | grep -v Warn
