"""DRAM traffic per launch of the fused MLP kernel from an ncu CSV log of the bench command:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_mlp_tc6 -c 400 \
        --csv --log-file gpurun_out/tc6_dram.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline
    python tools/ncu_dram_per_launch.py gpurun_out/tc6_dram.csv profiles/rNN_ncu_k_mlp_tc6_dram.json "<command note>"

bench.py reads the newest profiles/r*_ncu_k_mlp_tc6_dram.json for `roofline.traffic` (cited with its sha256)."""
import collections
import csv
import json
import sys


def main(path, out, note=''):
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    per = collections.OrderedDict()
    for r in rows:
        if 'k_mlp_tc6' not in r['Kernel Name']:
            continue
        d = per.setdefault(r['ID'], {})
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit'].lower()
        scale = {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'nsecond': 1e-6, 'ms': 1.0, 'msecond': 1.0, 'second': 1e3}.get(unit, 1)
        d[r['Metric Name']] = v * scale
    launches = [d for d in per.values() if 'dram__bytes_read.sum' in d]
    tot = [d['dram__bytes_read.sum'] + d['dram__bytes_write.sum'] for d in launches]
    res = {'command': note, 'kernel': 'k_mlp_tc6', 'launches_captured': len(launches),
           'dram_bytes_per_launch': sum(tot) / max(len(tot), 1), 'dram_bytes_total': sum(tot),
           'gpu_time_ms_total_under_ncu': sum(d.get('gpu__time_duration.sum', 0.0) for d in launches),
           'per_launch_dram_bytes': [round(t) for t in tot]}
    json.dump(res, open(out, 'w'), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != 'per_launch_dram_bytes'}))


if __name__ == '__main__':
    main(*sys.argv[1:4])
