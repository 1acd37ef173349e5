O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q -k "radius or pipeline or fullsize or stage or random or dropin or exchange" > $O/pytest_r2ab.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_r2ab.log
run() { # tag env lanes
env $2 timeout 200 python bench.py --no-cpu-baseline --no-configs --no-e2e --frames-total 512 --steps 12 --lanes $3 > $O/bench_r2ab_$1.json 2> $O/bench_r2ab_$1.err
python -c "
import json; d=json.load(open('$O/bench_r2ab_$1.json')); ga=[k for k in d['kernels'] if k['kernel']=='k_grid_assign'][0]['us_per_launch']; print('$1', d['value'], round(d['ms_per_step']*1e3/d['config']['frames_per_step_per_gpu'],2), 'us/scan', 'p50', d['p50_latency_ms'], 'k_grid_assign', ga)"
}
for rep in 1 2; do
run celllist_$rep "X=1" 6
run nocelllist_$rep "APC_NO_CELL_LIST=1" 6
done
