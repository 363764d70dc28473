#!/bin/bash
# round 2, trip f: ncu --set full of k_clers_cta<4> on 16 x configs[1] meshes (1 CTA per SM) to see where a window step spends its time
set -u
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_clers_cta --launch-skip 3 -c 1 -f -o $O/r2f_clers_cta \
    python bench.py --batch 16 --distinct 4 --steps 1 --warmup 3 --no-cpu --no-e2e > $O/r2f_ncu_clers.log 2>&1
ncu -i $O/r2f_clers_cta.ncu-rep --page details > $O/r2f_clers_cta_details.txt 2>&1
tail -3 $O/r2f_ncu_clers.log | cut -c1-300
