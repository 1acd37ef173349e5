#!/bin/bash
# N = $1 with the round's final code (exchange fused), quick variant
N=$1; TAG=${2:-r2}
O=gpurun_out; mkdir -p $O
Q="--no-configs --no-cpu-baseline --steps 20 --warmup 5"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N $Q > $O/bench_${TAG}_n$N.json 2> $O/bench_${TAG}_n$N.err; echo "n$N rc=$?"
python -c "
import json; d=json.load(open('$O/bench_${TAG}_n$N.json')); print('n$N', d['value'], d['ms_per_step'], (d['e2e'] or {}).get('value'), d['exchange_check'] and d['exchange_check']['mismatches'])"
