#!/bin/bash
# round 2, trip p: parity after the probe change; ncu --set full of the TIMED configuration (256 x configs[1]) for the four
# heaviest kernels; launch list of one step
set -u
O=gpurun_out
timeout 2000 python -X faulthandler -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/r2p_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $O/r2p_pytest_gpu.txt
grep -v "^  File" $O/r2p_pytest_gpu.txt | tail -4
for w in c4 tarta; do
  extra=""; [ $w = c4 ] && extra="--distinct 64"
  timeout 300 python bench.py --workload $w --steps 3 --no-cpu --no-e2e --no-shard --no-secondary $extra > $O/r2p_bench_$w.json 2> $O/r2p_bench_$w.err
  python -c "import json;d=json.loads(open('$O/r2p_bench_$w.json').read().strip().splitlines()[-1]);print('$w', round(d['ms_per_step'],3), round(d['value']), {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})" || tail -3 $O/r2p_bench_$w.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_clers_cta|k_delta_mesh_seg|k_adj_build|k_normal_estimate|k_unpack_chain' --launch-skip 15 -c 5 -f -o $O/r2p_c2_256 \
    python bench.py --distinct 16 --steps 1 --warmup 3 --no-cpu --no-e2e --no-shard --no-secondary > $O/r2p_ncu.log 2>&1
ncu -i $O/r2p_c2_256.ncu-rep --page details > $O/r2p_c2_256_details.txt 2>&1
grep -E "^  [a-z_:]+.*\(|^    Duration|DRAM Throughput|Executed Ipc Active|Achieved Occupancy|^    Registers|dram__bytes" $O/r2p_c2_256_details.txt | head -60
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2p_launches_c2.csv python bench.py --distinct 16 --steps 2 --warmup 1 --no-cpu --no-e2e --no-shard --no-secondary > $O/r2p_launches.log 2>&1
tail -30 $O/r2p_launches_c2.csv | cut -c1-200
