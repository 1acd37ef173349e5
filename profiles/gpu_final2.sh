#!/bin/bash
# Final visit of the round: parity tests, smoke, the driver's bench command + reference arm, ncu launch list.
TAG=${1:-rX}
O=gpurun_out; mkdir -p $O
timeout 500 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_$TAG.log
timeout 100 python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$TAG.json 2>> $O/bench_$TAG.err; echo "ref rc=$?"
SMALL="--steps 2 --warmup 3 --frames-total 12 --no-cpu-baseline --no-configs --no-e2e"
timeout 200 python bench.py $SMALL > $O/bench_small_$TAG.json 2>> $O/bench_$TAG.err &&
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file $O/launches_$TAG.csv \
    python bench.py $SMALL > $O/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
python -c "
import json; d=json.load(open('$O/bench_$TAG.json')); print(d['value'], d['ms_per_step'], 'p50', d['p50_latency_ms'], d['p50_latency_low_latency_lane_ms'], 'e2e', d['e2e']['value'], d['p50_latency_e2e_ms'], 'roof', d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'], [ (c, d['configs'][c]['value'], d['configs'][c]['p50_latency_ms']) for c in ('C1','C3','C4')])"
