O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_r2g.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_r2g.log
Q="--no-configs --no-cpu-baseline --no-e2e --frames-total 256 --steps 30"
python bench.py $Q > $O/bench_r2g_base.json 2>> $O/bench_r2g.err
APC_RS_CH=20 python bench.py $Q > $O/bench_r2g_ch20.json 2>> $O/bench_r2g.err
APC_VOX_ITEMS=1 python bench.py $Q > $O/bench_r2g_vox1.json 2>> $O/bench_r2g.err
APC_RADIUS_SPLIT=0 python bench.py $Q > $O/bench_r2g_nosplit.json 2>> $O/bench_r2g.err
APC_RS_CH=20 APC_VOX_ITEMS=1 APC_RADIUS_SPLIT=0 python bench.py $Q > $O/bench_r2g_old.json 2>> $O/bench_r2g.err
for f in base ch20 vox1 nosplit old; do python -c "
import json; d=json.load(open('$O/bench_r2g_$f.json')); print('$f', d['value'], round(d['ms_per_step']*1e3/d['config']['frames_per_step_per_gpu'],2), 'us/scan', d['kernels_per_scan'], d['p50_latency_ms'], [(k['kernel'],k['us_per_launch']) for k in d['kernels'][:14]])"; done
bash profiles/rs_trace.sh r2g_ch10
