"""Dynamic per-source-line attribution: joins an ncu source-page CSV (SASS, executed counts, stall
samples) with nvdisasm --print-line-info of the cubin by instruction offset.
  python scripts/ncu_lines.py rep.ncu-rep cubin kernel_substr [n]"""
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
inst = collections.Counter(); samp = collections.Counter(); fp64 = collections.Counter()
tot_i = tot_s = 0
for r in rows:
    off = int(r["Address"], 16) - base
    ln = off2line.get(off, "?")
    n = int(r["Instructions Executed"] or 0); s = int(r["# Samples"] or 0)
    inst[ln] += n; samp[ln] += s; tot_i += n; tot_s += s
    op = r["Source"].split()[1] if r["Source"].strip().startswith("@") else r["Source"].split()[0]
    if op.split(".")[0] in ("DADD", "DMUL", "DFMA", "DSETP"):
        fp64[ln] += n
print(f"total warp-instr {tot_i}  samples {tot_s}  fp64 instr {sum(fp64.values())}")
for ln, s in samp.most_common(topn):
    print(f"{ln:32s} samples {100*s/tot_s:5.1f}%  instr {100*inst[ln]/tot_i:5.1f}%  fp64 {fp64[ln]/1e6:7.1f}M")
