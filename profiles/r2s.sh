O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_r2s.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_r2s.log
timeout 120 python profiles/stage_times.py 8 > $O/stage_times_r2s.log 2>&1; echo "stage_times rc=$?"; cat $O/stage_times_r2s.log
for L in 8 12 16; do
timeout 200 python bench.py --no-cpu-baseline --no-configs --no-e2e --frames-total 256 --steps 20 --lanes $L > $O/bench_r2s_l$L.json 2> $O/bench_r2s_l$L.err; echo "bench lanes=$L rc=$?"
python -c "
import json; d=json.load(open('$O/bench_r2s_l$L.json')); print('lanes=$L', d['value'], round(d['ms_per_step']*1e3/d['config']['frames_per_step_per_gpu'],2), 'us/scan', 'p50', d['p50_latency_ms'])"
done
