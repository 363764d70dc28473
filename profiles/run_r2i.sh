#!/bin/bash
# round 2, trip i: ncu --set full of k_clers_cta<4> on configs[4] (one 10 M-vertex mesh: pure window throughput)
set -u
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_clers_cta --launch-skip 1 -c 1 -f -o $O/r2i_clers_c5 \
    python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/r2i_ncu_c5.log 2>&1
ncu -i $O/r2i_clers_c5.ncu-rep --page details > $O/r2i_clers_c5_details.txt 2>&1
tail -2 $O/r2i_ncu_c5.log | cut -c1-200
