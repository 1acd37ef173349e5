#!/bin/bash
# usage: gpurun --gpus 8 --timeout 600 -- 'bash profiles/e2e_scale_probe.sh TAG'
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/topo_$TAG.txt 2>&1
lscpu > $O/lscpu_$TAG.txt 2>&1
free -g >> $O/lscpu_$TAG.txt 2>&1
for N in 8 4 2 1; do
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
      profiles/e2e_scale_probe.py $O/e2e_scale_probe_${TAG}_n$N.json > $O/e2e_scale_probe_${TAG}_n$N.log 2>&1
  echo "N=$N rc=$?"
done
