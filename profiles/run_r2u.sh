#!/bin/bash
# round 2, trip u: adjacency (faces only) + boundary scan + non-critical delta on the side stream beside the position delta; parity; overlap A/B
set -u
O=gpurun_out
timeout 2000 python -X faulthandler -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2u_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $O/r2u_pytest_gpu.txt
grep -v "^  File" $O/r2u_pytest_gpu.txt | tail -8
for ov in 0 2 3; do
CORTO_OVERLAP=$ov timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-secondary --no-shard --distinct 16 > $O/r2u_bench_ov$ov.json 2> $O/r2u_bench_ov$ov.err
python -c "import json;d=json.loads(open('$O/r2u_bench_ov$ov.json').read().strip().splitlines()[-1]);print('c2 overlap $ov', round(d['ms_per_step'],3), round(d['value']))" || tail -3 $O/r2u_bench_ov$ov.err
done
for w in c4 c5; do for ov in 0 2 3; do
extra=""; [ $w = c4 ] && extra="--distinct 64"
CORTO_OVERLAP=$ov timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu --no-secondary --no-shard $extra > $O/r2u_bench_${w}_ov$ov.json 2> $O/r2u_bench_${w}_ov$ov.err
python -c "import json;d=json.loads(open('$O/r2u_bench_${w}_ov$ov.json').read().strip().splitlines()[-1]);print('$w overlap $ov', round(d['ms_per_step'],3), round(d['value']))" || tail -3 $O/r2u_bench_${w}_ov$ov.err
done; done
