"""Which source lines make up the HOT code footprint (static instructions that carry 99.9 % of the dynamic
warp instructions)?   python scripts/ncu_hotlines.py rep.ncu-rep cubin kernel_substr [n]"""
import collections, csv, io, re, subprocess, sys
rep, cubin, sub = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
base = int(rows[0]["Address"], 16)
txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
funcs = re.split(r'\n\s*//-+ \.text\.(\S+)', txt)
off2line = {}
for i in range(1, len(funcs), 2):
    if sub not in funcs[i]:
        continue
    cur = None
    for l in funcs[i + 1].splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = f"{m.group(1).split('/')[-1]}:{m.group(2)}"
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/', l)
        if m:
            off2line[int(m.group(1), 16)] = cur
    break
ex = [(int(r["Instructions Executed"] or 0), int(r["Address"], 16) - base) for r in rows]
tot = sum(e for e, _ in ex)
ex.sort(reverse=True)
acc = 0
hot = collections.Counter(); dyn = collections.Counter()
n = 0
for e, off in ex:
    if acc >= 0.999 * tot:
        break
    acc += e; n += 1
    hot[off2line.get(off)] += 1
    dyn[off2line.get(off)] += e
print(f"hot set: {n} instructions = {n * 16 / 1024:.1f} KB")
for line, c in hot.most_common(topn):
    print(f"{line:32s} {c:5d} instr  {c * 16:6d} B   dyn {100 * dyn[line] / tot:5.1f}%")
