"""Summarise ncu outputs into small text tables for profiles/ (the .ncu-rep files stay in gpurun_out/, untracked).

    python tools/ncu_summary.py launches gpurun_out/r1_launches.csv            > profiles/rNN_launches.md
    python tools/ncu_summary.py full gpurun_out/r1_prof_conv.ncu-rep           > profiles/rNN_conv_full.md
"""
import collections
import csv
import io
import subprocess
import sys

FULL = ['Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum',
        'sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.sum',
        'sm__ops_path_tensor_op_utcqmma_src_fp4_fp6_fp8_dst_fp32_sparsity_off.sum',
        'smsp__sass_inst_executed_op_utcmma.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max']


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = row['Kernel Name'].split('(')[0].replace('void ', '')
        v = float(row['Metric Value'].replace(',', ''))
        v = v / 1000 if row['Metric Unit'] == 'ns' else (v * 1000 if row['Metric Unit'] == 'ms' else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    print('ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold-cache: compare SHARES)\n')
    print('%d launches, %.1f us total\n' % (n, tot))
    print('| kernel | launches | us | share |\n|---|---|---|---|')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| %s | %d | %.1f | %.1f%% |' % (k, a[0], a[1], 100 * a[1] / tot))


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [(w, hdr.index(w)) for w in FULL if w in hdr]
    kn = hdr.index('Kernel Name')
    print('ncu --set full --clock-control none, per launch (%s)\n' % path)
    for r in rows[2:]:
        print('### %s' % r[kn][:100])
        for w, i in cols:
            if r[i] not in ('', 'no data'):
                print('- %s = %s %s' % (w, r[i], units[i]))
        print()


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
