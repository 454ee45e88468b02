timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],2),'fps', round(d['ms_per_step'],2), {k:round(x,2) for k,x in d['roofline']['stage_ms_per_step'].items()}, 'mlp', round(d['roofline']['kernel_ms_per_step'],2))"
timeout 600 python tools/bench_ground.py 512 2 2>&1 | tail -1 | cut -c1-400
