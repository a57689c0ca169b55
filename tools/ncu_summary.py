"""Summarise an .ncu-rep of the force kernel: key launch metrics, stall mix, opcode mix, hottest SASS lines.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv, collections, subprocess, sys
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
want=["gpu__time_duration.sum","dram__bytes_read.sum ","dram__bytes_write.sum ","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread ","smsp__issue_active.avg.pct","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active","sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","smsp__average_warps_issue_stalled","smsp__thread_inst_executed_per_inst_executed","launch__waves","sm__inst_executed.sum ","launch__occupancy_limit"]
for h,u,v in zip(hdr,units,vals):
    if any(w in h+" " for w in want):
        try:
            if 'issue_stalled' in h and float(v)<0.05: continue
        except: pass
        print(h,u,v)
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))[1:]
hdr=rows[0]; ci={h:i for i,h in enumerate(hdr)}
data=rows[1:]
tot=0; byop=collections.Counter(); samp=collections.Counter(); top=[]
stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for r in data:
    if len(r)<len(hdr): continue
    t=r[ci['Source']].split()
    op=t[1] if t and t[0].startswith('@') else (t[0] if t else '?')
    op=op.split('.')[0]
    n=int(r[ci['Instructions Executed']]); s=int(r[ci['# Samples']])
    tot+=n; byop[op]+=n; samp[op]+=s
    top.append((s,r[ci['Address']][-5:],r[ci['Source']].strip()[:50],n,{c[6:]:int(r[ci[c]]) for c in stall_cols if int(r[ci[c]])>s*0.3 and s>50}))
print("total inst",tot)
for op,n in byop.most_common(16): print("%-10s %6.2f%% inst %6.2f%% samples"%(op,100*n/tot,100*samp[op]/sum(samp.values())))
top.sort(reverse=True)
for t in top[:14]: print(t)
