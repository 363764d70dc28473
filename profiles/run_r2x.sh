#!/bin/bash
# round 2, trip x: A/B of the stage-overlap bits on one box (configs[1], 20 steps each, twice)
set -u
O=gpurun_out
for rep in 1 2; do for ov in 0 2 3 7; do
CORTO_OVERLAP=$ov timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-secondary --no-shard --distinct 16 > $O/r2x_bench_ov$ov.json 2> $O/r2x_bench_ov$ov.err
python -c "import json;d=json.loads(open('$O/r2x_bench_ov$ov.json').read().strip().splitlines()[-1]);print('c2 overlap $ov', round(d['ms_per_step'],3), round(d['value']))" || tail -3 $O/r2x_bench_ov$ov.err
done; done
