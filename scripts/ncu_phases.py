"""Instruction share / lane utilisation by phase of the BVH traversal kernels, from an .ncu-rep holding one or more
launches:  python scripts/ncu_phases.py rep.ncu-rep cubin"""
import collections, csv, io, re, subprocess, sys
rep, cubin = sys.argv[1:3]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-kernel-base", "function"], capture_output=True, text=True).stdout
blocks = []; cur = None; name = None
for l in out.splitlines():
    if l.startswith('"Kernel Name"'):
        name = l.split(",")[1].strip('"'); continue
    if l.startswith('"Address"'):
        cur = [l]; blocks.append((name, cur)); continue
    if cur is not None and l.startswith('"0x'):
        cur.append(l)
txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
funcs = re.split(r'\n\s*//-+ \.text\.(\S+)', txt)
sizes = {}
maps = {}
for i in range(1, len(funcs), 2):
    m = {}; c = None
    for l in funcs[i + 1].splitlines():
        mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if mm: c = (mm.group(1).split('/')[-1], int(mm.group(2))); continue
        mm = re.match(r'\s+/\*([0-9a-f]{4,5})\*/', l)
        if mm: m[int(mm.group(1), 16)] = c
    maps[funcs[i]] = m
def classify(ln):
    if not ln: return "other"
    f, n = ln
    if f == "nt_bvh_trace.cuh":
        if 45 <= n <= 98: return "query_start"
        if 100 <= n <= 109: return "pop"
        if 111 <= n <= 143: return "inner_step"
        if 145 <= n <= 175: return "leaf_step"
        return "state machine / shading"
    if f == "nt_wavefront.cuh": return "wavefront task fetch / setup / retire"
    if f == "nt_trace.cuh": return "primitive tests, math"
    return "other"
for bi, (name, b) in enumerate(blocks):
    rows = list(csv.DictReader(io.StringIO("\n".join(b))))
    n_static = len(rows)
    cands = [k for k, m in maps.items() if name in k and len(m) == n_static]
    mp = maps[cands[0]] if cands else {}
    base = int(rows[0]["Address"], 16)
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for r in rows:
        ln = mp.get(int(r["Address"], 16) - base)
        e = int(r["Instructions Executed"] or 0); t = int(r["Thread Instructions Executed"] or 0); sm = int(r["# Samples"] or 0)
        k = classify(ln)
        agg[k][0] += e; agg[k][1] += t; agg[k][2] += sm
    tot = sum(v[0] for v in agg.values()); ts = sum(v[2] for v in agg.values())
    print(f"launch {bi}: {cands[0] if cands else name}  warp instr {tot}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"    {k:40s} instr {100 * v[0] / tot:5.1f}%  samples {100 * v[2] / max(ts, 1):5.1f}%  avg threads {v[1] / max(v[0], 1):5.1f}")
