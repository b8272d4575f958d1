import csv, io, subprocess, sys, re, bisect
rep = sys.argv[1]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur=None;h2=None;agg={}
for r in rows:
    if len(r)==2 and r[0] in ("File Path","File Name"): cur=r[1]; continue
    if len(r)==2: continue
    if r[0]=="Line No": h2=r; continue
    if r[2] not in ("-",""): continue
    try: ln=int(r[0])
    except: continue
    def g(name):
        try: return int(r[h2.index(name)])
        except: return 0
    a=agg.setdefault((cur,ln),[0,0,0,r[1]])
    a[0]+=g("# Samples"); a[1]+=g("Instructions Executed"); a[2]+=g("Thread Instructions Executed")
files={}
for (f,ln),a in agg.items(): files.setdefault(f,{})[ln]=a
tot_s=sum(a[0] for a in agg.values()) or 1; tot_i=sum(a[1] for a in agg.values()) or 1
res={}
for f,lines in files.items():
    # function starts from the real file (current tree; line numbers may have drifted a little)
    starts=[]
    try:
        src=open(f).read().split('\n')
    except Exception:
        src=[]
    for i,txt in enumerate(src,1):
        m=re.match(r'^(?:static\s+)?(?:EUCL_HD|__device__|__global__|__host__|inline)\b[^;]*?\b([A-Za-z_0-9]+)\s*\(',txt)
        if m: starts.append((i,m.group(1)))
    for ln,a in lines.items():
        i=bisect.bisect_right([s[0] for s in starts],ln)-1
        fn=starts[i][1] if i>=0 else '?'
        k=(f.split('/')[-1],fn)
        r_=res.setdefault(k,[0,0,0]); r_[0]+=a[0]; r_[1]+=a[1]; r_[2]+=a[2]
print("samples%% inst%% lanes  file:function   (total warp inst %d)"%tot_i)
for k,v in sorted(res.items(),key=lambda kv:-kv[1][1])[:45]:
    print(f"{v[0]/tot_s*100:6.1f} {v[1]/tot_i*100:6.1f} {v[2]/max(v[1],1)/32:5.2f}  {k[0]}:{k[1]}")
