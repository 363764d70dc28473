#!/bin/bash
# profiles/collect_r2.sh — the round-2 evidence, ONE gpurun call on one B200:
#   parity (pytest -m gpu), smoke(), the default bench line (configs[1] + e2e + cpu_baseline + shard + secondary), the reference arm,
#   the launch list of the same command, compute-sanitizer memcheck over smoke().
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 600 > $O/r2_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $O/r2_pytest_gpu.txt
grep -v "^  File" $O/r2_pytest_gpu.txt | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.txt 2>&1; tail -2 $O/r2_smoke.txt
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2_bench_c2_reference.json 2> $O/r2_bench_c2_reference.err
tail -c 600 $O/r2_bench_c2_reference.json
timeout 1200 python bench.py > $O/r2_bench_c2.json 2> $O/r2_bench_c2.err
tail -3 $O/r2_bench_c2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_c2.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e'] and round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4), 'step_achieved', round(d['roofline']['step_achieved']))
print({k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})
print('cpu', d['cpu_baseline']); print('shard', d['shard']); print('secondary', d['secondary']); print('clocks', d['clocks'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-shard --no-secondary > $O/r2_launches.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_memcheck.log 2>&1; tail -3 $O/r2_memcheck.log
