#!/bin/bash
# round 2, trip j: prev-chain fast path also for chains in the reach-back store (configs[4]); parity + workloads
set -u
O=gpurun_out
timeout 2000 python -X faulthandler -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2j_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $O/r2j_pytest_gpu.txt
grep -v "^  File" $O/r2j_pytest_gpu.txt | tail -5
for w in c2 c5 c3; do
  extra=""; [ $w = c2 ] && extra="--distinct 16"; [ $w = c3 ] && extra="--distinct 16"
  timeout 300 python bench.py --workload $w --steps 3 --no-cpu --no-e2e $extra > $O/r2j_bench_$w.json 2> $O/r2j_bench_$w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-300:], open(f.replace('.json','.err')).read()[-500:])
PY
