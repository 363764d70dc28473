#!/bin/bash
# round 2, trip t: stage overlap A/B on configs[1]: 0 none, 1 unpack beside CLERS, 2 non-critical delta beside the normal estimation, 3 both
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "alternative" -p no:cacheprovider 2>&1 | tail -3
for ov in 0 1 2 3; do
CORTO_OVERLAP=$ov timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-secondary --no-shard --distinct 16 > $O/r2t_bench_ov$ov.json 2> $O/r2t_bench_ov$ov.err
python -c "import json;d=json.loads(open('$O/r2t_bench_ov$ov.json').read().strip().splitlines()[-1]);print('overlap $ov', round(d['ms_per_step'],3), round(d['value']))" || tail -3 $O/r2t_bench_ov$ov.err
done
for ov in 0 3; do
CORTO_OVERLAP=$ov timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-e2e --no-cpu --no-secondary --no-shard --distinct 64 > $O/r2t_bench_c4_ov$ov.json 2> $O/r2t_bench_c4_ov$ov.err
python -c "import json;d=json.loads(open('$O/r2t_bench_c4_ov$ov.json').read().strip().splitlines()[-1]);print('c4 overlap $ov', round(d['ms_per_step'],3), round(d['value']))" || tail -3 $O/r2t_bench_c4_ov$ov.err
done
