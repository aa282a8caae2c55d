#!/bin/bash
# Round-2 session h (2 GPUs): bench under torchrun (NCCL), sharded graph self-check.
O=gpurun_out/r2h; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --cells-total 400000 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_graph_check.py > $O/dist_graph.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 --cpu-seconds 20 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
tail -c 800 $O/bench_n2.err; head -c 2500 $O/bench_n2.json; echo; tail -5 $O/dist_graph.log; head -c 300 $O/bench_ref_n2.json
