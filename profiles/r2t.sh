O=gpurun_out; mkdir -p $O
for S in 2 4; do for L in 6 8; do
APC_HASH_SLOTS_PER_POINT=$S timeout 200 python bench.py --no-cpu-baseline --no-configs --no-e2e --frames-total 256 --steps 20 --lanes $L > $O/bench_r2t_s${S}_l$L.json 2> $O/bench_r2t_s${S}_l$L.err; echo "bench slots=$S lanes=$L rc=$?"
python -c "
import json; d=json.load(open('$O/bench_r2t_s${S}_l$L.json')); print('slots=$S lanes=$L', d['value'], round(d['ms_per_step']*1e3/d['config']['frames_per_step_per_gpu'],2), 'us/scan', 'p50', d['p50_latency_ms'], 'tp-lane p50', d['p50_latency_throughput_lane_ms'])"
done; done
