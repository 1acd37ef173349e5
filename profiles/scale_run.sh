#!/bin/bash
# The C5 scaling run: bench.py at N = 8 / 4 / 1 on one 8-GPU box (+ N = 8 with NVLS multicast stores when
# the fabric offers them, + N = 8 without any exchange for attribution).
# usage: gpurun --gpus 8 --timeout 900 -- 'bash profiles/scale_run.sh TAG'
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
Q="--no-configs --no-cpu-baseline --steps 20 --warmup 5"
run() {  # N port suffix extra-env
  env $4 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 \
      bench.py --gpus $1 $Q > $O/bench_${TAG}_$3.json 2> $O/bench_${TAG}_$3.err
  echo "$3 rc=$?"
}
python profiles/multicast_probe.py 8 > $O/multicast_probe_$TAG.txt 2>&1; tail -2 $O/multicast_probe_$TAG.txt
run 8 29601 n8 "APC_MIRROR=p2p"
if grep -q "multicast_ptr [1-9]" $O/multicast_probe_$TAG.txt; then run 8 29602 n8_multicast "APC_MIRROR=multicast"; fi
run 8 29603 n8_noexchange "APC_GATHER=none"
run 4 29604 n4 "APC_MIRROR=p2p"
python bench.py $Q > $O/bench_${TAG}_n1.json 2> $O/bench_${TAG}_n1.err; echo "n1 rc=$?"
for f in n1 n4 n8 n8_multicast n8_noexchange; do [ -s $O/bench_${TAG}_$f.json ] && python -c "
import json; d=json.load(open('$O/bench_${TAG}_$f.json')); print('$f', d['value'], d['ms_per_step'], (d['e2e'] or {}).get('value'), d['exchange_check'] and d['exchange_check']['mismatches'], d['config']['multi_gpu'][:90])"; done
grep -h "self-check" $O/bench_${TAG}_n8.err | head -3
