#!/bin/bash
# round 2, trip h: shared-state races of k_clers_cta fixed (second barrier after the dispatch reads, pop state in registers, one
# slot of slack in the symbol ring): parity, synccheck + racecheck on a 320-mesh configs[3] batch (R = 2048 path), workloads
set -u
O=gpurun_out
mkdir -p $O
timeout 2000 python -X faulthandler -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2h_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $O/r2h_pytest_gpu.txt
grep -v "^  File" $O/r2h_pytest_gpu.txt | tail -30
timeout 900 compute-sanitizer --tool synccheck --print-limit 10 python profiles/c4_batch_320.py 320 > $O/r2h_synccheck.txt 2>&1
grep -v "Host Frame" $O/r2h_synccheck.txt | tail -12 | cut -c1-250
timeout 1500 compute-sanitizer --tool racecheck --print-limit 30 python profiles/c4_batch_320.py 320 > $O/r2h_racecheck.txt 2>&1
grep -v "Host Frame" $O/r2h_racecheck.txt | tail -30 | cut -c1-250
for w in c2 c5 tarta c4; do
  extra=""; [ $w = c2 ] && extra="--distinct 16"; [ $w = c4 ] && extra="--distinct 64"
  timeout 300 python bench.py --workload $w --steps 3 --no-cpu --no-e2e $extra > $O/r2h_bench_$w.json 2> $O/r2h_bench_$w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2h_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-300:], open(f.replace('.json','.err')).read()[-500:])
PY
