O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q -k "radius or pipeline or fullsize or stage or random or dropin" > $O/pytest_r2z.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_r2z.log
for v in 2,0 2,8 1,8 3,8 2,4 2,32; do
APC_RADIUS_COOP=$v timeout 120 python profiles/stage_times.py 8 > $O/stage_times_r2z_$v.log 2>&1; echo "coop=$v $(head -1 $O/stage_times_r2z_$v.log) $(grep k_radius_query $O/stage_times_r2z_$v.log)"
APC_RADIUS_COOP=$v timeout 200 python bench.py --no-cpu-baseline --no-configs --no-e2e --frames-total 256 --steps 20 > $O/bench_r2z_$v.json 2> $O/bench_r2z_$v.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$O/bench_r2z_$v.json')); print('coop=$v', d['value'], round(d['ms_per_step']*1e3/d['config']['frames_per_step_per_gpu'],2), 'us/scan', 'p50', d['p50_latency_ms'])"
done
