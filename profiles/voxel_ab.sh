#!/bin/bash
# hash-vs-sort evidence: timings (plain run), then per-kernel DRAM bytes + L2 hit rates (ncu)
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
python profiles/voxel_ab.py $O/voxel_ab_$TAG.json > $O/voxel_ab_$TAG.log 2>&1 && \
VOXEL_AB_REPS=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:'k_voxel|k_sort' --csv --log-file $O/voxel_ab_ncu_$TAG.csv \
    python profiles/voxel_ab.py > $O/voxel_ab_ncu_$TAG.log 2>&1
echo "voxel_ab rc=$?"; cat $O/voxel_ab_$TAG.log | tail -4
