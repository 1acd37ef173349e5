O=gpurun_out; mkdir -p $O
run() { # tag env lanes
env $2 timeout 200 python bench.py --no-cpu-baseline --no-configs --no-e2e --frames-total 512 --steps 12 --lanes $3 > $O/bench_r2ab_$1.json 2> $O/bench_r2ab_$1.err
python -c "
import json; d=json.load(open('$O/bench_r2ab_$1.json')); ga=[k for k in d['kernels'] if k['kernel']=='k_radius_query'][0]['us_per_launch']; print('$1', d['value'], round(d['ms_per_step']*1e3/d['config']['frames_per_step_per_gpu'],2), 'us/scan', 'p50', d['p50_latency_ms'], 'k_radius_query', ga)"
}
run blk128_1 "APC_RADIUS_BLOCK=128" 6
run blk64_1 "APC_RADIUS_BLOCK=64" 6
run blk256_1 "APC_RADIUS_BLOCK=256" 6
run blk128_2 "APC_RADIUS_BLOCK=128" 6
run blk64_2 "APC_RADIUS_BLOCK=64" 6
