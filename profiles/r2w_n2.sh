# two GPUs: peer-slab test + the driver's N = 2 bench command (exchange fused, self-check)
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_exchange.py -m gpu -x -q > $O/pytest_r2w.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_r2w.log
Q="--no-configs --no-cpu-baseline --steps 10 --warmup 3"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 $Q > $O/bench_r2w_n2.json 2> $O/bench_r2w_n2.err; echo "n2 rc=$?"
tail -3 $O/bench_r2w_n2.err
python -c "
import json; d=json.load(open('$O/bench_r2w_n2.json')); print('n2', d['value'], d['ms_per_step'], d['e2e'], d['exchange_check'], d['config']['multi_gpu'][:120])"
