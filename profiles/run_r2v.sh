#!/bin/bash
# round 2, trip w: normal estimation split into the valence <= 8 launch and the high-valence launch; parity; workloads
set -u
O=gpurun_out
timeout 2000 python -X faulthandler -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2w_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $O/r2w_pytest_gpu.txt
grep -v "^  File" $O/r2w_pytest_gpu.txt | tail -6
for w in c2 c4 c3 tarta c5; do
  extra=""; [ $w = c2 ] && extra="--distinct 16"; [ $w = c3 ] && extra="--distinct 16"; [ $w = c4 ] && extra="--distinct 64"
  timeout 300 python bench.py --workload $w --steps 5 --no-cpu --no-e2e --no-shard --no-secondary $extra > $O/r2w_bench_$w.json 2> $O/r2w_bench_$w.err
  python -c "import json;d=json.loads(open('$O/r2w_bench_$w.json').read().strip().splitlines()[-1]);print('$w', round(d['ms_per_step'],3), round(d['value']), {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})" || tail -3 $O/r2w_bench_$w.err
done
