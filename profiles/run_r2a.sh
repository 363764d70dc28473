#!/bin/bash
# first GPU trip of round 2: parity of the new CTA-wide CLERS kernel + the new config tests, then stage times per workload
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2a_pytest_gpu.txt
cat $O/r2a_pytest_gpu.txt
for wl in c2 tarta c5; do
  timeout 300 python bench.py --workload $wl --steps 5 --no-cpu --no-e2e --distinct 16 > $O/r2a_bench_$wl.json 2> $O/r2a_bench_$wl.err
  CORTO_CLERS=3 timeout 300 python bench.py --workload $wl --steps 5 --no-cpu --no-e2e --distinct 16 > $O/r2a_bench_${wl}_lf.json 2> $O/r2a_bench_${wl}_lf.err
done
CORTO_RUNMIN=2 timeout 300 python bench.py --workload tarta --steps 5 --no-cpu --no-e2e > $O/r2a_bench_tarta_rm2.json 2>&1
CORTO_RUNMIN=8 timeout 300 python bench.py --workload tarta --steps 5 --no-cpu --no-e2e > $O/r2a_bench_tarta_rm8.json 2>&1
timeout 300 python bench.py --workload c4 --steps 5 --no-cpu --no-e2e --distinct 64 > $O/r2a_bench_c4.json 2> $O/r2a_bench_c4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-300:])
PY
