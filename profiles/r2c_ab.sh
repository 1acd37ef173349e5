O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_r2c.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_r2c.log
python bench.py > $O/bench_r2c.json 2> $O/bench_r2c.err; echo "bench rc=$?"
Q="--no-configs --no-cpu-baseline --no-e2e"
python bench.py $Q --frames-total 64 --steps 60 > $O/bench_r2c_f64.json 2>> $O/bench_r2c.err
APC_RS_HALF_CTAS_PER_SM=4 python bench.py $Q > $O/bench_r2c_rs4.json 2>> $O/bench_r2c.err
python bench.py $Q --lanes 4 > $O/bench_r2c_l4.json 2>> $O/bench_r2c.err
python bench.py $Q --lanes 16 > $O/bench_r2c_l16.json 2>> $O/bench_r2c.err
APC_HASH_SLOTS_PER_POINT=2 python bench.py $Q > $O/bench_r2c_h2.json 2>> $O/bench_r2c.err
for f in r2c r2c_f64 r2c_rs4 r2c_l4 r2c_l16 r2c_h2; do python -c "
import json; d=json.load(open('$O/bench_$f.json')); print('$f', d['value'], d['ms_per_step'], d['p50_latency_ms'], d['pipeline_roofline'].get('steady_state_counters'))"; done
