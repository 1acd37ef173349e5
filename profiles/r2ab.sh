O=gpurun_out; mkdir -p $O
run() { # tag env lanes
env $2 timeout 200 python bench.py --no-cpu-baseline --no-configs --no-e2e --frames-total 512 --steps 12 --lanes $3 > $O/bench_r2ab_$1.json 2> $O/bench_r2ab_$1.err
python -c "
import json; d=json.load(open('$O/bench_r2ab_$1.json')); print('$1', d['value'], round(d['ms_per_step']*1e3/d['config']['frames_per_step_per_gpu'],2), 'us/scan', 'p50', d['p50_latency_ms'], 'kernels', d['kernels_per_scan'])"
}
for rep in 1 2; do
run base_l6_$rep "X=1" 6
run streamin_l6_$rep "APC_STREAM_IN=1" 6
run streamin_l8_$rep "APC_STREAM_IN=1" 8
done
