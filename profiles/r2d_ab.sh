O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_r2d.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_r2d.log
Q="--no-configs --no-cpu-baseline --no-e2e --frames-total 256 --steps 30"
python bench.py $Q > $O/bench_r2d_base.json 2>> $O/bench_r2d.err
APC_DUMMY_KERNELS=4 python bench.py $Q > $O/bench_r2d_d4.json 2>> $O/bench_r2d.err
APC_DUMMY_KERNELS=13 python bench.py $Q > $O/bench_r2d_d13.json 2>> $O/bench_r2d.err
python bench.py $Q --lanes 2 > $O/bench_r2d_l2.json 2>> $O/bench_r2d.err
python bench.py $Q --lanes 1 > $O/bench_r2d_l1.json 2>> $O/bench_r2d.err
for f in base d4 d13 l2 l1; do python -c "
import json; d=json.load(open('$O/bench_r2d_$f.json')); print('$f', d['value'], round(d['ms_per_step']*1e3/d['config']['frames_per_step_per_gpu'],2), 'us/scan', d['kernels_per_scan'], d['p50_latency_ms'])"; done
bash profiles/voxel_ab.sh r2d
