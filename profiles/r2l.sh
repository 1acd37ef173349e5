O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_r2l.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_r2l.log
python profiles/voxel_ab.py $O/voxel_ab_r2l.json 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-e2e --frames-total 256 --steps 30 > $O/bench_r2l.json 2> $O/bench_r2l.err
python -c "
import json; d=json.load(open('$O/bench_r2l.json')); print(d['value'], round(d['ms_per_step']*1e3/d['config']['frames_per_step_per_gpu'],2), 'us/scan', d['p50_latency_ms'], [(k['kernel'],k['us_per_launch']) for k in d['kernels'][:13]]); print({k:(v.get('value'),v.get('ms_per_scan'),v.get('p50_latency_ms')) for k,v in d['configs'].items()})"
