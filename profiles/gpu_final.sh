#!/bin/bash
# Round-end visit: parity tests, smoke, the driver's bench command, reference arm, then (each after its plain
# run exited 0) the ncu launch list, one full-set capture of an eager scan, and the steady-state DRAM capture.
# usage: gpurun --timeout 1700 -- 'bash profiles/gpu_final.sh TAG'
TAG=${1:-rX}
O=gpurun_out; mkdir -p $O
timeout 500 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_$TAG.log
timeout 100 python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$TAG.json 2>> $O/bench_$TAG.err; echo "ref rc=$?"
SMALL="--steps 2 --warmup 3 --frames-total 16 --no-cpu-baseline --no-configs --no-e2e"
timeout 200 python bench.py $SMALL > $O/bench_small_$TAG.json 2>> $O/bench_$TAG.err &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches_$TAG.csv \
    python bench.py $SMALL > $O/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
timeout 100 python profiles/run_pipeline.py 3 > $O/plain_$TAG.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^k_ --launch-skip 24 -c 14 -f -o $O/prof_$TAG \
    python profiles/run_pipeline.py 3 > $O/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
ARGS="--frames-total 64 --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-e2e"
timeout 500 ncu --cache-control none --clock-control none --metrics $M -k regex:^k_ --launch-skip 1500 -c 704 --csv \
    --log-file $O/steady_kernels_$TAG.csv python bench.py $ARGS > $O/steady_kernels_$TAG.log 2>&1
echo "steady capture rc=$?"
tail -c 1500 $O/bench_$TAG.json
