#!/bin/bash
# DRAM traffic of the pipeline AS IT RUNS (8 lanes, warm L2 shared by the lanes), not of a cold eager scan:
#  1. range replay over the timed region of a short bench run (all lanes concurrent, no cache flush);
#  2. fall-back: kernel replay without cache control (kernels serialised, but the L2 keeps whatever the
#     other lanes' kernels left in it).
# usage: gpurun --timeout 1200 -- 'bash profiles/steady_traffic.sh TAG'
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
ARGS="--frames-total 64 --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-e2e"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum
APC_PROFILE_RANGE=1 timeout 600 ncu --replay-mode range --profile-from-start off --cache-control none --clock-control none \
    --metrics $M --csv --log-file $O/steady_range_$TAG.csv python bench.py $ARGS > $O/steady_range_$TAG.log 2>&1
echo "range replay rc=$?"
timeout 600 ncu --cache-control none --clock-control none --metrics $M -k regex:^k_ --launch-skip 1700 -c 832 --csv \
    --log-file $O/steady_kernels_$TAG.csv python bench.py $ARGS > $O/steady_kernels_$TAG.log 2>&1
echo "kernel replay rc=$?"
tail -5 $O/steady_range_$TAG.csv | cut -c1-400
