"""Opcode mix + hottest instructions of an ncu source-page CSV: python scripts/ncu_opmix.py rep.ncu-rep [n]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
ops = collections.Counter(); thr = collections.Counter(); samp = collections.Counter()
tot = 0
for r in rows:
    src = r["Source"].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    n = int(r["Instructions Executed"] or 0)
    ops[op] += n; tot += n
    thr[op] += int(r["Thread Instructions Executed"] or 0)
    samp[op] += int(r["# Samples"] or 0)
print("total warp instr", tot, "static instr", len(rows))
for op, n in ops.most_common(topn):
    print(f"{op:10s} {n:12d} {100*n/tot:5.1f}%  avg_thr {thr[op]/max(n,1):5.1f}  samples {samp[op]}")
print("--- hottest by samples")
rows.sort(key=lambda r: -int(r["# Samples"] or 0))
for r in rows[:topn]:
    print(r["# Samples"], r["Instructions Executed"], r["Source"].strip()[:90], "| wait", r.get("stall_wait"), "ssb", r.get("stall_short_sb"), "bar", r.get("stall_barrier"), "math", r.get("stall_math"))
