"""Basic-block view of a scripts/ncu_sass_dump.py listing: runs of instructions with the same execution count, with their
share of the dynamic warp instructions and of the stall samples.   python scripts/ncu_regions.py sass.txt [min_percent]"""
import sys, collections
rows=[]
for l in open(sys.argv[1]):
    p=l.split(None,6)
    rows.append((int(p[0],16),int(p[1]),float(p[2]),int(p[3]),int(p[4]),p[5],p[6].strip()))
tot=sum(r[1] for r in rows); samp=sum(r[3] for r in rows)
print("total dyn",tot,"samples",samp)
# group contiguous instructions with similar exec count (within 2%) 
groups=[]; cur=[rows[0]]
for r in rows[1:]:
    a=cur[-1][1]; b=r[1]
    if (a==b) or (a>0 and b>0 and abs(a-b)/max(a,b)<0.03): cur.append(r)
    else: groups.append(cur); cur=[r]
groups.append(cur)
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.4
for g in groups:
    dyn=sum(r[1] for r in g); s=sum(r[3] for r in g); ni=sum(r[4] for r in g)
    if dyn/tot*100<thr: continue
    lines=collections.Counter(r[5] for r in g)
    ops=collections.Counter(r[6].split()[0] if not r[6].startswith('@') else r[6].split()[1] for r in g)
    d=sum(1 for r in g if r[6].split()[0 if not r[6].startswith('@') else 1].startswith(('DADD','DMUL','DFMA','DSETP')))
    print(f"{g[0][0]:6x}-{g[-1][0]:6x} n={len(g):4d} exec={g[0][1]:9d} dyn={dyn/tot*100:5.2f}% samp={s/samp*100:5.2f}% noinst={ni:5d} thr={g[0][2]:4.1f} fp64={d:3d} lines={[k.split(':')[1] for k,_ in lines.most_common(5)]}")
