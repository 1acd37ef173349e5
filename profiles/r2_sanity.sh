timeout 200 python bench.py --no-cpu-baseline --no-configs --frames-total 256 --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['p50_latency_ms'], d['e2e']['value'], d['pipeline_roofline'])"
