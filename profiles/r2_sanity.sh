O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_sanity.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_sanity.log
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python bench.py --no-cpu-baseline --no-configs --frames-total 256 --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['p50_latency_ms'], d['e2e']['value'], d['kernels_per_scan'], d['config']['lanes'])"
