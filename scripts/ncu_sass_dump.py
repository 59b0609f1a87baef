"""Per-instruction dump of one kernel from an .ncu-rep: offset, warp executions, average active threads, stall samples,
stall_no_inst samples, source line (joined with nvdisasm --print-line-info of the cubin), SASS.
  python scripts/ncu_sass_dump.py rep.ncu-rep kernels.cubin kernel_substr > sass.txt   (cubin: cuobjdump -xelf all lib.so)"""
import csv, io, re, subprocess, sys
rep, cubin, sub = sys.argv[1:4]  # python scripts/ncu_sass_dump.py rep.ncu-rep kernels.cubin kernel_substr > sass.txt
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
print(list(rows[0].keys()), file=sys.stderr)
base = int(rows[0]["Address"], 16)
txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
funcs = re.split(r'\n\s*//-+ \.text\.(\S+)', txt)
off2line = {}
for i in range(1, len(funcs), 2):
    if sub not in funcs[i]: continue
    cur = None
    for l in funcs[i + 1].splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/', l)
        if m: off2line[int(m.group(1), 16)] = cur
for r in rows:
    off = int(r["Address"], 16) - base
    ln = off2line.get(off)
    ex = r.get("Instructions Executed") or "0"
    th = r.get("Thread Instructions Executed") or "0"
    sm = r.get("# Samples") or r.get("Warp Stall Sampling (All Samples)") or "0"
    ni = r.get("stall_no_inst") or "0"
    print(f"{off:6x} {int(ex):10d} {int(th)/max(int(ex),1):5.1f} {sm:>6} {ni:>5} {ln[0] if ln else '?'}:{ln[1] if ln else 0:<5} {r['Source'].strip()}")
