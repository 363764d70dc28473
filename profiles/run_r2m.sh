#!/bin/bash
# round 2, trip m: does the attribute unpack beside the CLERS automaton (CORTO_OVERLAP=1) pay now that CLERS is one CTA per mesh?
set -u
O=gpurun_out
for ov in 0 1; do
CORTO_OVERLAP=$ov timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-secondary --no-shard --distinct 16 > $O/r2m_bench_ov$ov.json 2> $O/r2m_bench_ov$ov.err
python -c "import json;d=json.loads(open('$O/r2m_bench_ov$ov.json').read().strip().splitlines()[-1]);print('overlap $ov', d['ms_per_step'], d['value'])"
done
