O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_r2u.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_r2u.log
timeout 100 python __graft_entry__.py smoke > $O/smoke_r2u.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_r2u.log
for V in fold nofold; do
if [ $V = nofold ]; then export APC_NO_FOLD=1; fi
timeout 200 python bench.py --no-cpu-baseline --no-configs --no-e2e --frames-total 256 --steps 20 > $O/bench_r2u_$V.json 2> $O/bench_r2u_$V.err; echo "bench $V rc=$?"
python -c "
import json; d=json.load(open('$O/bench_r2u_$V.json')); print('$V', d['value'], round(d['ms_per_step']*1e3/d['config']['frames_per_step_per_gpu'],2), 'us/scan', 'p50', d['p50_latency_ms'], 'tp-lane', d['p50_latency_throughput_lane_ms'], 'kernels', d['kernels_per_scan'])"
done
