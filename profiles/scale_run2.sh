#!/bin/bash
# N = 8 on one 8-GPU box with the round's final code: exchange fused (NVLS multicast when offered) and no exchange.
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
Q="--no-configs --no-cpu-baseline --steps 20 --warmup 5"
run() {  # N port suffix extra-env
  env $4 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 \
      bench.py --gpus $1 $Q > $O/bench_${TAG}_$3.json 2> $O/bench_${TAG}_$3.err
  echo "$3 rc=$?"
}
run 8 29601 n8 "APC_MIRROR=auto"
run 8 29603 n8_noexchange "APC_GATHER=none"
for f in n8 n8_noexchange; do [ -s $O/bench_${TAG}_$f.json ] && python -c "
import json; d=json.load(open('$O/bench_${TAG}_$f.json')); print('$f', d['value'], d['ms_per_step'], (d['e2e'] or {}).get('value'), d['exchange_check'] and d['exchange_check']['mismatches'], d['config']['multi_gpu'][:160])"; done
grep -h "self-check" $O/bench_${TAG}_n8.err | head -3
