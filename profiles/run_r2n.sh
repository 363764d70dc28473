#!/bin/bash
# round 2, trip n: hybrid CLERS launch (regular meshes on k_clers_cta, irregular ones deferred to k_clers_lf), parity + tarta / c2;
# attribute unpack beside the automaton (CORTO_OVERLAP=1) A/B
set -u
O=gpurun_out
timeout 2000 python -X faulthandler -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2n_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $O/r2n_pytest_gpu.txt
grep -v "^  File" $O/r2n_pytest_gpu.txt | tail -8
for w in c2 tarta; do
  extra=""; [ $w = c2 ] && extra="--distinct 16"
  timeout 300 python bench.py --workload $w --steps 3 --no-cpu --no-e2e --no-shard --no-secondary $extra > $O/r2n_bench_$w.json 2> $O/r2n_bench_$w.err
  python -c "import json;d=json.loads(open('$O/r2n_bench_$w.json').read().strip().splitlines()[-1]);print('$w', d['ms_per_step'], d['value'], {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})" || tail -3 $O/r2n_bench_$w.err
done
for ov in 0 1; do
CORTO_OVERLAP=$ov timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-secondary --no-shard --distinct 16 > $O/r2n_bench_ov$ov.json 2> $O/r2n_bench_ov$ov.err
python -c "import json;d=json.loads(open('$O/r2n_bench_ov$ov.json').read().strip().splitlines()[-1]);print('overlap $ov', d['ms_per_step'], d['value'])"
done
