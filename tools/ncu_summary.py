"""Summarise an .ncu-rep here (no GPU needed): headline counters + SASS opcode mix + hottest SASS segments.
usage: python tools/ncu_summary.py report.ncu-rep [launch-index]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else '0'
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--launch-skip', skip, '--launch-count', '1'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__cycles_active.avg', 'smsp__issue_active.avg.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'launch__registers_per_thread', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size']
for i, h in enumerate(hdr):
    if h in keys:
        print(f'{h:75s} {units[i]:12s} {r[i]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--launch-skip', skip,
                      '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
iS, iE, iW = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = [(x[iS], int(x[iE] or 0), int(x[iW] or 0), x) for x in rows[2:] if len(x) > iE]
tot, tots = sum(d[1] for d in data), max(1, sum(d[2] for d in data))
print(f'\nSASS: {len(data)} instructions, {tot} executed (warp level), {tots} stall samples')
st = Counter()
for d in data:
    for i in stall_cols:
        st[hdr[i]] += int(d[3][i] or 0)
print('stall reasons:', ', '.join(f'{k[6:]} {100 * v / tots:.1f}%' for k, v in st.most_common(8)))
c, cs = Counter(), Counter()
for s, e, w, _ in data:
    parts = s.split()
    op = (parts[1] if s.startswith('@') else parts[0]).split('.')[0]
    c[op] += e
    cs[op] += w
print('opcode mix:')
for op, e in c.most_common(16):
    print(f'  {op:10s} {e:10d} {100 * e / tot:5.1f}%   samples {100 * cs[op] / tots:5.1f}%')
segs = []
for i, (s, e, w, _) in enumerate(data):
    if segs and segs[-1][2] == e:
        segs[-1][1] = i; segs[-1][3] += e; segs[-1][4] += w
    else:
        segs.append([i, i, e, e, w])
print('segments (consecutive SASS with equal execution count) above 1.5% of instructions or 2% of samples:')
for a, b, e, t, w in segs:
    if t > 0.015 * tot or w > 0.02 * tots:
        ops = ' '.join((d[0].split()[1] if d[0].startswith('@') else d[0].split()[0]).split('.')[0] for d in data[a:min(b + 1, a + 12)])
        print(f'  [{a:4d}-{b:4d}] n={b - a + 1:4d} exec={e:8d} inst {100 * t / tot:5.1f}% samples {100 * w / tots:5.1f}% :: {ops}')
