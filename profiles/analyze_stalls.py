import csv,re,collections,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
r=list(csv.reader(raw.splitlines()))
d=dict(zip(r[0],r[2]))
g=lambda k: float(d[k])
print(f"t={g('gpu__time_duration.sum'):.2f} inst={g('smsp__inst_executed.sum')/1e6:.0f}M issue={g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}% warps/smsp={g('smsp__warps_active.avg.per_cycle_active'):.2f} thr/inst={g('smsp__thread_inst_executed_per_inst_executed.ratio')}")
st=[]
for k in d:
    if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
        st.append((round(float(d[k]),2),k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
print(sorted(st,reverse=True)[:8])
for k in ['l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']:
    if k in d: print(k,d[k])
syms=[]
for l in open(sys.argv[2] if len(sys.argv) > 2 else '/tmp/syms.txt'):   # 'value size name' lines from readelf -sW of the kernel's cubin
    v,sz,name=l.split()
    short=name.split('$')[-1]
    short=re.sub(r'_ZN8x265b200\d*me_thread_only','',short)[:40] if '$' in name else 'KERNEL'
    syms.append((int(v,16),int(sz),short))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
base=int(data[0][ix['Address']],16)
def f(x):
    try: return float(x)
    except: return 0.0
agg=collections.defaultdict(lambda:[0]*6)
for r in data:
    off=int(r[ix['Address']],16)-base
    name='KERNEL'
    for v,sz,n in syms:
        if v<=off<v+sz: name=n
    a=agg[name]
    a[0]+=f(r[ix['# Samples']]); a[1]+=f(r[ix['Instructions Executed']]); a[2]+=f(r[ix['stall_long_sb']]); a[3]+=f(r[ix['stall_wait']]); a[4]+=f(r[ix['stall_no_inst']]); a[5]+=f(r[ix['stall_short_sb']])
tot=sum(a[0] for a in agg.values())
for n,a in sorted(agg.items(),key=lambda x:-x[1][0])[:9]:
    print(f"{n:42s} samples {a[0]/tot*100:5.1f}%  inst {a[1]/1e6:7.1f}M  long_sb {a[2]/tot*100:5.1f}% wait {a[3]/tot*100:5.1f}% no_inst {a[4]/tot*100:5.1f}% short {a[5]/tot*100:4.1f}%")
