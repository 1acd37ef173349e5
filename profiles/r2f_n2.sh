O=gpurun_out; mkdir -p $O
python - <<'PY' > $O/multicast_probe_r2f.txt 2>&1
import os, torch, torch.distributed as dist, torch.multiprocessing as mp
def w(rank):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", RANK=str(rank), WORLD_SIZE="2")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    import torch.distributed._symmetric_memory as symm
    t = symm.empty((1024,), dtype=torch.float32, device=f"cuda:{rank}")
    h = symm.rendezvous(t, dist.group.WORLD)
    print(rank, "multicast_ptr", getattr(h, "multicast_ptr", None), "has_multicast", getattr(h, "has_multicast_support", None), flush=True)
    dist.destroy_process_group()
if __name__ == "__main__":
    mp.spawn(w, nprocs=2)
PY
cat $O/multicast_probe_r2f.txt | tail -4
python -m pytest tests/test_gpu_exchange.py tests/test_gpu_stages.py -m gpu -x -q > $O/pytest_r2f.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_r2f.log
Q="--no-configs --no-cpu-baseline"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 $Q > $O/bench_r2f_n2.json 2> $O/bench_r2f_n2.err; echo "n2 rc=$?"
tail -5 $O/bench_r2f_n2.err
APC_GATHER=none python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 $Q --no-e2e > $O/bench_r2f_n2_none.json 2>> $O/bench_r2f_n2.err; echo "n2 none rc=$?"
python bench.py $Q --no-e2e > $O/bench_r2f_n1.json 2>> $O/bench_r2f_n2.err
for f in n1 n2 n2_none; do python -c "
import json; d=json.load(open('$O/bench_r2f_$f.json')); print('$f', d['value'], d['ms_per_step'], d['e2e'], d['exchange_check'])"; done
