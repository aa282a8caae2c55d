set -x
timeout -s KILL 200 python -u -m pytest tests/test_infercnv_gpu.py -m gpu -x -q --timeout 100 --timeout-method=thread 2>&1 | tail -3
python bench.py --workload csr100 --steps 50 --no-cpu-baseline > gpurun_out/bench_csr100.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"smooth_kernel|center_rows_kernel|colsum_dense_kernel|apply_threshold_kernel|gene_values_kernel|dense_to_csr_kernel" -o gpurun_out/step_r1b python tools/one_step.py 100000 100 > gpurun_out/one_step.log 2>&1
tail -c 400 gpurun_out/bench_csr100.json
