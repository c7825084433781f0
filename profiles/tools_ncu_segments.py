import csv,sys,subprocess
rep=sys.argv[1]; samples=float(sys.argv[2]) if len(sys.argv)>2 else 1e6; thr=float(sys.argv[3]) if len(sys.argv)>3 else 0.3
raw=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
h=rows[1]
isrc=h.index('Source'); ie=h.index('Instructions Executed'); it=h.index('Avg. Threads Executed'); ismp=h.index('# Samples')
data=rows[2:]
wd=samples*564/32
# group into contiguous segments with equal weight
seg=[]; 
for k,r in enumerate(data):
    c=int(r[ie])/wd
    seg.append((k,c,float(r[it]),int(r[ismp]),r[isrc].strip()))
tot=sum(s[1] for s in seg); print("total",tot)
# print compact: runs of similar weight
i=0
while i<len(seg):
    j=i
    while j+1<len(seg) and abs(seg[j+1][1]-seg[i][1])<0.02*max(seg[i][1],0.05) and abs(seg[j+1][2]-seg[i][2])<1.5: j+=1
    w=sum(s[1] for s in seg[i:j+1]); smp=sum(s[3] for s in seg[i:j+1])
    if w>=thr:
        ops={}
        for s in seg[i:j+1]:
            o=s[4].split()[0] if not s[4].startswith('@') else s[4].split()[1]
            o=o.split('.')[0]; ops[o]=ops.get(o,0)+1
        top=sorted(ops.items(),key=lambda x:-x[1])[:8]
        print(f"[{i:4d}-{j:4d}] n={j-i+1:3d} w/instr={seg[i][1]:6.2f} thr={seg[i][2]:4.1f} total={w:7.2f} samples={smp:7d}  {top}")
    i=j+1
