#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, then (each only after its plain run exited 0) the
# ncu launch list of the bench command and one full-set capture of every kernel of one eager scan.
# usage: gpurun --timeout 1500 -- 'bash profiles/gpu_round.sh TAG [quick]'
TAG=${1:-rX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
if [ "$2" != "quick" ]; then
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$TAG.json 2>> $O/bench_$TAG.err; echo "ref rc=$?"
SMALL="--steps 2 --warmup 3 --frames-total 16 --no-cpu-baseline --no-configs --no-e2e"
python bench.py $SMALL > $O/bench_small_$TAG.json 2>> $O/bench_$TAG.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches_$TAG.csv \
    python bench.py $SMALL > $O/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
python profiles/run_pipeline.py 3 > $O/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:^k_ --launch-skip 28 -c 20 -f -o $O/prof_$TAG \
    python profiles/run_pipeline.py 3 > $O/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
fi
tail -3 $O/pytest_$TAG.log; tail -c 3000 $O/bench_$TAG.json
