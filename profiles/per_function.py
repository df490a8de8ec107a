#!/usr/bin/env python
"""Per-function view of an ncu capture of a kernel that calls __noinline__ device functions:
   python profiles/per_function.py REPORT.ncu-rep SYMS.txt
SYMS.txt = 'value size name' lines (hex value) of the kernel's cubin:  cuobjdump -xelf ... ; readelf -sW x.cubin | awk '$4=="FUNC"{print $2,$3,$8}'"""
import csv, collections, re, subprocess, sys
rep, symf = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines())); d = dict(zip(r[0], r[2])); g = lambda k: float(d[k])
print(f"t={g('gpu__time_duration.sum'):.3f} {r[1][r[0].index('gpu__time_duration.sum')]} inst={g('smsp__inst_executed.sum')/1e6:.0f}M issue={g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}% "
      f"warps_active={g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}% thr/inst={d['smsp__thread_inst_executed_per_inst_executed.ratio']} regs={d['launch__registers_per_thread']}")
st = sorted(((round(float(d[k]), 2), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for k in d
             if 'issue_stalled' in k and k.endswith('per_issue_active.ratio')), reverse=True)[:8]
print(st)
for k in ('l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'):
    print(k, d.get(k))
syms = []
for l in open(symf):
    v, sz, name = l.split()
    syms.append((int(v, 16), int(sz), re.sub(r'^\d+', '', name.split('$')[-1])[:34]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); hdr = rows[1]; data = rows[2:]; ix = {h: i for i, h in enumerate(hdr)}
f = lambda x: float(x) if x.replace('.', '', 1).isdigit() else 0.0
base = int(data[0][ix['Address']], 16)
agg = collections.defaultdict(lambda: [0] * 6); tot = 0
for rw in data:
    off = int(rw[ix['Address']], 16) - base
    name = 'KERNEL body'
    for v, sz, n in syms:
        if v <= off < v + sz: name = n
    a = agg[name]; smp = f(rw[ix['# Samples']])
    a[0] += smp; a[1] += f(rw[ix['Instructions Executed']]); a[2] += f(rw[ix['stall_long_sb']]); a[3] += f(rw[ix['stall_wait']]); a[4] += f(rw[ix['stall_no_inst']]); a[5] += f(rw[ix['stall_short_sb']])
    tot += smp
for n, a in sorted(agg.items(), key=lambda x: -x[1][0])[:14]:
    print(f"{n:36s} samples {a[0]/tot*100:5.1f}%  inst {a[1]/1e6:8.1f}M  long_sb {a[2]/tot*100:5.1f}% wait {a[3]/tot*100:5.1f}% no_inst {a[4]/tot*100:5.1f}% short_sb {a[5]/tot*100:4.1f}%")
