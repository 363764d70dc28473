#!/bin/bash
# second GPU trip of round 2: parity first, then an ncu capture of k_clers_cta (16 x c2) to see where its window steps spend their time
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_robust.py -m gpu -q 2>&1 | tail -25 > $O/r2b_pytest_gpu.txt
cat $O/r2b_pytest_gpu.txt
timeout 300 python bench.py --workload c2 --steps 5 --no-cpu --no-e2e --distinct 16 > $O/r2b_bench_c2.json 2> $O/r2b_bench_c2.err
timeout 300 python bench.py --workload c5 --steps 3 --no-cpu --no-e2e > $O/r2b_bench_c5.json 2> $O/r2b_bench_c5.err
timeout 300 python bench.py --workload tarta --steps 3 --no-cpu --no-e2e > $O/r2b_bench_tarta.json 2> $O/r2b_bench_tarta.err
timeout 300 python bench.py --workload c4 --steps 3 --no-cpu --no-e2e --distinct 64 > $O/r2b_bench_c4.json 2> $O/r2b_bench_c4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_clers_cta --launch-skip 3 -c 1 -f -o $O/r2b_clers_cta \
    python bench.py --batch 16 --distinct 4 --steps 1 --warmup 3 --no-cpu --no-e2e > $O/r2b_ncu_clers.log 2>&1
ncu -i $O/r2b_clers_cta.ncu-rep --page details > $O/r2b_clers_cta_details.txt 2>&1
ncu -i $O/r2b_clers_cta.ncu-rep --page source --csv > $O/r2b_clers_cta_source.csv 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-300:])
PY
