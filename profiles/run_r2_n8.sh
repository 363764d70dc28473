#!/bin/bash
# round 2: the default bench line on 8 GPUs of one box (NCCL over NVLink): weak-scaling value + e2e, and the shard block (strong scaling)
set -u
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r2_bench_c2_n8.json 2> $O/r2_bench_c2_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_c2_n8.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e'] and round(d['e2e']['value']))
print('shard', json.dumps(d['shard'], indent=1))
PY
tail -3 $O/r2_bench_c2_n8.err
