#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:knn_mma -c 1 --csv --log-file $O/knn1m.csv python tools/knn_one.py 1000000 50 > $O/ncu.log 2>&1
grep -v "^==" $O/knn1m.csv | cut -d, -f5,13- | tail -9
