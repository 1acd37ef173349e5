# A/B of the radius grid: cells of 2r + pruned 8-cell walk (default) against cells of r + 27-cell walk
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q > $O/pytest_r2p.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_r2p.log
for v in 2 1; do
APC_RADIUS_CELL=$v timeout 120 python profiles/stage_times.py 8 > $O/stage_times_r2p_cell$v.log 2>&1; echo "stage_times cell=$v rc=$?"; cat $O/stage_times_r2p_cell$v.log
APC_RADIUS_CELL=$v timeout 200 python bench.py --no-cpu-baseline --no-configs --no-e2e --frames-total 256 --steps 20 > $O/bench_r2p_cell$v.json 2> $O/bench_r2p_cell$v.err; echo "bench cell=$v rc=$?"
python -c "
import json; d=json.load(open('$O/bench_r2p_cell$v.json')); print('cell=$v', d['value'], round(d['ms_per_step']*1e3/d['config']['frames_per_step_per_gpu'],2), 'us/scan', 'p50', d['p50_latency_ms'])"
done
