"""Static SASS instruction count per source line of one kernel: python scripts/sass_lines.py cubin kernel_substr [n]"""
import re, collections, subprocess, sys
cubin, sub = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
funcs = re.split(r'\n\s*//-+ \.text\.(\S+)', txt)
for i in range(1, len(funcs), 2):
    name, body = funcs[i], funcs[i + 1]
    if sub not in name:
        continue
    cur = None; cnt = collections.Counter(); total = 0
    for line in body.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,5}\*/', line):
            cnt[cur] += 1; total += 1
    print(name, total, "instr", total * 16 / 1024, "KB")
    for (f, l), n in cnt.most_common(topn):
        print(f"{f}:{l}  {n}")
