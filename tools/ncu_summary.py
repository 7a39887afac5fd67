#!/usr/bin/env python
"""Summarise ncu outputs into small text/JSON files for profiles/ (the .ncu-rep files themselves stay in gpurun_out/).
  python tools/ncu_summary.py launches gpurun_out/r1_launches.csv profiles/r1_launches_summary.txt
  python tools/ncu_summary.py full gpurun_out/r1_quad_stream.ncu-rep profiles/r1_quad_stream_ncu.txt [profiles/quad_traffic.json BATCH]
"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    iK, iV, iI = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('ID')
    per = OrderedDict()
    order = []
    for r in rows[1:]:
        name = r[iK].split('(')[0].replace('void ', '')
        ns = float(r[iV].replace(',', ''))
        per.setdefault(name, []).append(ns)
        order.append((int(r[iI]), name, ns))
    tot = sum(sum(v) for v in per.values())
    with open(dst, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n')
        f.write('# source: %s ; %d launches, %.1f us total\n' % (src, len(order), tot/1e3))
        f.write('%-44s %8s %12s %10s %8s\n' % ('kernel', 'launches', 'total_us', 'mean_us', 'share'))
        for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write('%-44s %8d %12.1f %10.2f %7.1f%%\n' % (name[:44], len(v), sum(v)/1e3, sum(v)/len(v)/1e3, 100*sum(v)/tot))
        # share inside one timed step: the config-3 sweep pair (weights_sweep -> rows) when the run has it, else the
        # config-2 triple (weights, stream, epilogue) with the longest stream kernel
        if any('hm_quad_rows' in o[1] for o in order):
            seq = [o for o in order if 'hm_weights_sweep' in o[1] or 'hm_quad_rows' in o[1]]
            cands = [i for i in range(len(seq)-1) if 'hm_weights_sweep' in seq[i][1] and 'hm_quad_rows' in seq[i+1][1]]
            if not cands:
                return print(open(dst).read())
            best = max(cands, key=lambda i: seq[i+1][2])
            step, label = seq[best:best+2], 'one sweep sub-batch (weights -> row-per-lane sweep kernel)'
        else:
            trip = [o for o in order if any(t in o[1] for t in ('hm_weights', 'hm_quad_stream', 'hm_epilogue'))]
            cands = [i for i in range(len(trip)-2) if 'hm_weights' in trip[i][1] and 'hm_quad_stream' in trip[i+1][1]
                     and 'hm_epilogue' in trip[i+2][1]]
            best = max(cands, key=lambda i: trip[i+1][2])
            step, label = trip[best:best+3], 'one batch step (weights -> stream -> epilogue)'
        st = sum(o[2] for o in step)
        f.write('\n# %s: %.1f us\n' % (label, st/1e3))
        for _, name, ns in step:
            f.write('%-44s %10.2f us %7.1f%%\n' % (name[:44], ns/1e3, 100*ns/st))
    print(open(dst).read())


KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'sm__cycles_elapsed.max']


def full(src, dst, traffic_json=None, batch=None):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = ['# ncu --set full --clock-control none --import-source on ; source: '+src]
    traffic = []
    for r in rows[2:]:
        out.append('## '+r[hdr.index('Kernel Name')])
        for key in KEYS:
            if key in hdr:
                out.append('  %-72s %s %s' % (key, r[hdr.index(key)], units[hdr.index(key)]))
        rd, wr = r[hdr.index('dram__bytes_read.sum')], r[hdr.index('dram__bytes_write.sum')]
        scale = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.}
        traffic.append(float(rd)*scale[units[hdr.index('dram__bytes_read.sum')]]
                       + float(wr)*scale[units[hdr.index('dram__bytes_write.sum')]])
    sass = subprocess.run(['ncu', '-i', src, '--page', 'source', '--csv', '--print-source', 'sass'],
                          capture_output=True, text=True).stdout
    ops = {}
    for line in sass.splitlines():
        for op in ('UBLKCP', 'SYNCS', 'DMMA', 'DFMA', 'DMUL', 'DADD', 'SHFL', 'LDS', 'LDG', 'STG', 'BAR', 'UTMALDG'):
            if (' '+op) in line or (','+op) in line or ('"'+op) in line:
                ops[op] = ops.get(op, 0)+1
    out.append('## SASS mnemonics present (static count over the profiled kernels): '+json.dumps(ops))
    open(dst, 'w').write('\n'.join(out)+'\n')
    print('\n'.join(out))
    if traffic_json:
        json.dump(dict(kernel='hm_quad_stream_kernel', batch=int(batch), dram_bytes_per_launch=sum(traffic)/len(traffic),
                       source=src), open(traffic_json, 'w'))


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3])
    else:
        full(*sys.argv[2:])
