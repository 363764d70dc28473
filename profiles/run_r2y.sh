#!/bin/bash
# round 2, trip y: Tunstall decode split (CLERS streams first, attribute streams on the side stream); parity; overlap A/B
set -u
O=gpurun_out
timeout 600 python -X faulthandler -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 300 > $O/r2y_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $O/r2y_pytest_gpu.txt
grep -v "^  File" $O/r2y_pytest_gpu.txt | tail -6
for ov in 2 3; do
CORTO_OVERLAP=$ov timeout 120 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-secondary --no-shard --distinct 16 > $O/r2y_bench_ov$ov.json 2> $O/r2y_bench_ov$ov.err
python -c "import json;d=json.loads(open('$O/r2y_bench_ov$ov.json').read().strip().splitlines()[-1]);print('c2 overlap $ov', round(d['ms_per_step'],3), round(d['value']))" || tail -3 $O/r2y_bench_ov$ov.err
done
for ov in 2 3; do
CORTO_OVERLAP=$ov timeout 120 python bench.py --workload c4 --steps 5 --warmup 3 --no-e2e --no-cpu --no-secondary --no-shard --distinct 64 > $O/r2y_bench_c4_ov$ov.json 2> $O/r2y_bench_c4_ov$ov.err
python -c "import json;d=json.loads(open('$O/r2y_bench_c4_ov$ov.json').read().strip().splitlines()[-1]);print('c4 overlap $ov', round(d['ms_per_step'],3), round(d['value']))" || tail -3 $O/r2y_bench_c4_ov$ov.err
done
