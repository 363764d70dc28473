#!/bin/bash
# round 2, trip k (2 GPUs): the default bench line at N=1 (with shard + secondary blocks) and the same at N=2 over NCCL
set -u
O=gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2k_bench_n1.json 2> $O/r2k_bench_n1.err
tail -c 3000 $O/r2k_bench_n1.json; tail -5 $O/r2k_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2k_bench_n2.json 2> $O/r2k_bench_n2.err
tail -c 2500 $O/r2k_bench_n2.json; tail -5 $O/r2k_bench_n2.err
