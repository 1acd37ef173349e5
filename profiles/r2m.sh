O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_r2m.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_r2m.log
python __graft_entry__.py smoke > $O/smoke_r2m.log 2>&1; echo "smoke rc=$?"
python bench.py --no-cpu-baseline --frames-total 256 --steps 30 > $O/bench_r2m.json 2> $O/bench_r2m.err; echo "bench rc=$?"; tail -3 $O/bench_r2m.err
python -c "
import json; d=json.load(open('$O/bench_r2m.json')); print(d['value'], round(d['ms_per_step']*1e3/d['config']['frames_per_step_per_gpu'],2), 'us/scan', 'p50', d['p50_latency_ms'], 'p99', d['p99_latency_ms'], 'tp-lane', d['p50_latency_throughput_lane_ms'], 'e2e p50', d['p50_latency_e2e_ms'], d['e2e']['value'])"
