timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final2.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/bench_final2.json')); print(round(d['value'],2),'fps', round(d['ms_per_step'],2), d['e2e']['value'], d['clocks'], {k:round(x,2) for k,x in d['roofline']['stage_ms_per_step'].items()}, 'mlp', round(d['roofline']['kernel_ms_per_step'],2), d['roofline']['frac'])"
