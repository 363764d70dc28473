#!/bin/bash
# profiles/collect_r1.sh — the commands behind the round-1 files in profiles/ (run on the GPU box from the repo root:
#   gpurun --timeout 1500 -- 'bash profiles/collect_r1.sh').  Outputs land in gpurun_out/ and are copied here by hand.
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -2 > $O/r1_pytest_gpu.txt
timeout 400 python bench.py > $O/r1_bench_c2.json 2> $O/r1_bench_c2.err
timeout 400 python bench.py --impl reference > $O/r1_bench_c2_reference.json 2> $O/r1_bench_c2_reference.err
timeout 300 python bench.py --workload c3 --steps 5 > $O/r1_bench_c3.json 2> $O/r1_bench_c3.err
timeout 300 python bench.py --workload c4 --steps 5 --distinct 64 --no-cpu --no-e2e > $O/r1_bench_c4.json 2> $O/r1_bench_c4.err
# launch list (cold, serialised: shares only)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1_launches_c2.csv \
    python bench.py --steps 2 --warmup 1 --distinct 32 --no-cpu --no-e2e > $O/r1_launches_c2.log 2>&1
# full capture of the dominant kernel on a 16-mesh batch of the same workload (~40 replays stay short)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_clers_lf --launch-skip 3 -c 1 -f -o $O/r1_clers_final \
    python bench.py --batch 16 --distinct 4 --steps 1 --warmup 3 --no-cpu --no-e2e > $O/r1_ncu_clers_final.log 2>&1
# memcheck over the smoke decode (three fixtures through the whole kernel chain)
timeout 240 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $O/r1_memcheck.log 2>&1
tail -3 $O/r1_pytest_gpu.txt $O/r1_memcheck.log
head -c 600 $O/r1_bench_c2.json
