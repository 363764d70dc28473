#!/bin/bash
# round 2, trip l (2 GPUs): device-arena batch parity, then the shard block at N=1 and N=2
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "from_device or batch_mixed" -p no:cacheprovider 2>&1 | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-secondary > $O/r2l_bench_n1.json 2> $O/r2l_bench_n1.err
python -c "import json;d=json.loads(open('$O/r2l_bench_n1.json').read().strip().splitlines()[-1]);print(json.dumps(d['shard'],indent=1))"; tail -3 $O/r2l_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu > $O/r2l_bench_n2.json 2> $O/r2l_bench_n2.err
python -c "import json;d=json.loads(open('$O/r2l_bench_n2.json').read().strip().splitlines()[-1]);print(json.dumps(d['shard'],indent=1))"; tail -3 $O/r2l_bench_n2.err
