"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.
  python tools/ncu_summarize.py launches gpurun_out/launches_v6.csv "<command line note>" > profiles/rNN_ncu_launch_list.txt
  python tools/ncu_summarize.py full gpurun_out/prof_tc6.ncu-rep "<note>" > profiles/rNN_ncu_<kernel>_summary.txt"""
import csv, subprocess, sys, collections, io

KEEP = ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct', 'launch__block_size', 'launch__grid_size', 'launch__cluster', 'launch__occupancy_limit',
        'launch__registers_per_thread', 'lts__throughput.avg.pct', 'lts__t_bytes.sum ', 'lts__t_sectors_op_read.sum ', 'sm__cycles_elapsed.max', 'sm__inst_executed_pipe_tensor',
        'sm__pipe_tensor', 'sm__throughput.avg.pct', 'sm__warps_active.avg.pct', 'smsp__inst_executed.sum ', 'sm__inst_executed_pipe_xu', 'sm__inst_executed_pipe_fma',
        'sm__inst_executed_pipe_alu', 'gpu__dram_throughput', 'l1tex__throughput', 'smsp__cycles_active.avg ', 'sm__pipe_shared_cycles_active', 'launch__shared_mem_per_block_dynamic',
        'smsp__warp_issue_stalled', 'smsp__average_warp', 'smsp__issue_active.avg.pct')


def launches(path, note):
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    agg = collections.OrderedDict()
    for r in rows:
        if r['Metric Name'] != 'gpu__time_duration.sum':
            continue
        k = r['Kernel Name'].split('(')[0][:60]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += float(r['Metric Value'].replace(',', '')) / 1e6
    tot = sum(a[1] for a in agg.values())
    print(note)
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{k:60s} n={n:4d} total {t:9.3f} ms ({100 * t / tot:5.1f}%) avg {1e3 * t / n:9.1f} us')
    print(f'total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches')


def full(path, note):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(note)
    for r in rows[2:]:
        print(f'--- launch id {r[0]}: {r[hdr.index("Kernel Name")][:50]} grid {r[hdr.index("Grid Size")]} block {r[hdr.index("Block Size")]}')
        for i, h in enumerate(hdr):
            if any(h.startswith(k.strip()) if k.endswith(' ') else h.startswith(k) for k in KEEP):
                print(f'{h} [{units[i]}] = {r[i]}')


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else '')
