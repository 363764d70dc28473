#!/bin/bash
# round 2, trip g: parity + workloads after adaptive window width (1/2/4 x 256 by the last run length), contiguous symbol ring, pop probing 8 flags before the wide scan
set -u
O=gpurun_out
mkdir -p $O
timeout 2000 python -X faulthandler -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2g_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $O/r2g_pytest_gpu.txt
grep -v "^  File" $O/r2g_pytest_gpu.txt | tail -30
timeout 300 python bench.py --workload c2 --steps 5 --no-cpu --no-e2e --distinct 16 > $O/r2g_bench_c2.json 2> $O/r2g_bench_c2.err
timeout 300 python bench.py --workload c5 --steps 3 --no-cpu --no-e2e > $O/r2g_bench_c5.json 2> $O/r2g_bench_c5.err
timeout 300 python bench.py --workload tarta --steps 3 --no-cpu --no-e2e > $O/r2g_bench_tarta.json 2> $O/r2g_bench_tarta.err
timeout 300 python bench.py --workload c4 --steps 3 --no-cpu --no-e2e --distinct 64 > $O/r2g_bench_c4.json 2> $O/r2g_bench_c4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2g_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-300:], open(f.replace('.json','.err')).read()[-500:])
PY
