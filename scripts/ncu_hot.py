"""Hot-code footprint of a kernel from an .ncu-rep: how many static instructions (x16 bytes) carry N% of the
dynamic warp instructions, and where stall_no_inst samples fall.   python scripts/ncu_hot.py rep.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
ex = [int(r["Instructions Executed"] or 0) for r in rows]
tot = sum(ex)
order = sorted(range(len(ex)), key=lambda i: -ex[i])
print(f"static instructions {len(ex)} = {len(ex) * 16 / 1024:.1f} KB, dynamic warp instructions {tot}")
acc = 0
marks = [0.5, 0.8, 0.9, 0.95, 0.99, 0.999]
mi = 0
for n, i in enumerate(order, 1):
    acc += ex[i]
    while mi < len(marks) and acc >= marks[mi] * tot:
        print(f"  {marks[mi] * 100:5.1f}% of dynamic instructions come from {n} static instructions = {n * 16 / 1024:.1f} KB")
        mi += 1
never = sum(1 for e in ex if e == 0)
print(f"  never executed: {never} instructions = {never * 16 / 1024:.1f} KB")
noi = [int(r.get("stall_no_inst") or 0) for r in rows]
print("stall_no_inst samples", sum(noi), "of", sum(int(r["# Samples"] or 0) for r in rows))
