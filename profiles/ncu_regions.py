import csv, subprocess, io, sys
rep=sys.argv[1]
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
cur=None;h2=None;agg={}
def I(x):
    try:return int(x)
    except: return 0
for r in rows:
    if len(r)>=2 and r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if len(r)>3 and r[0]=="Line No": h2=r; continue
    if h2 and len(r)==len(h2) and r[0]!="":
        d=dict(zip(h2,r))
        agg[(cur,int(r[0]))]=(I(d["Instructions Executed"]),I(d["# Samples"]),I(d.get("Thread Instructions Executed",0)))
tot=sum(v[0] for v in agg.values()); ts=sum(v[1] for v in agg.values())
regs=eval(sys.argv[2])
print("total warp inst %.3e"%tot)
for a,b,name in regs:
    i=sum(v[0] for k,v in agg.items() if k[0]=='bwb_lane.cuh' and a<=k[1]<b)
    t=sum(v[2] for k,v in agg.items() if k[0]=='bwb_lane.cuh' and a<=k[1]<b)
    s=sum(v[1] for k,v in agg.items() if k[0]=='bwb_lane.cuh' and a<=k[1]<b)
    print("%-34s %5.1f%% inst  lanes %4.1f  samples %5.1f%%"%(name,100*i/tot, t/max(i,1), 100*s/ts))
for f in sorted(set(k[0] for k in agg)):
    i=sum(v[0] for k,v in agg.items() if k[0]==f); t=sum(v[2] for k,v in agg.items() if k[0]==f)
    print(f, "%.1f%%"%(100*i/tot), "lanes %.1f"%(t/max(i,1)))
