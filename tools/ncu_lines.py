"""Join an `ncu --page source --csv` SASS listing with `nvdisasm -g -c` line info of the same kernel: per source line
stall samples and executed instructions.  Usage: ncu_lines.py <ncu_source.csv> <nvdisasm.sass> [top_n]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ins = [dict(zip(hdr, r)) for r in rows[2:] if len(r) >= 8]
# nvdisasm: "//## File "...", line N" comments precede instructions; instructions look like "/*0000*/  OPCODE ...;"
line = None
file = None
seq = []
for l in open(sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        file, line = m.group(1).split("/")[-1], int(m.group(2))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        seq.append((file, line, l.strip()))
print(len(ins), "ncu instructions;", len(seq), "nvdisasm instructions")
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
n = min(len(ins), len(seq))
for i in range(n):
    f, ln, txt = seq[i]
    d = ins[i]
    a = agg[(f, ln)]
    a[0] += int(float(d["# Samples"] or 0))
    a[1] += int(float(d["Instructions Executed"] or 0))
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print("samples", tot_s, "warp-instructions", tot_i)
srcs = {}
def text(f, ln):
    import glob, os
    if f not in srcs:
        c = glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dtqn_b200", "csrc", f))
        srcs[f] = open(c[0]).read().split("\n") if c else []
    return srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else ""
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:<5d} samples {a[0]:6d} ({100*a[0]/max(1,tot_s):4.1f}%)  inst {a[1]:9d} ({100*a[1]/max(1,tot_i):4.1f}%)  {text(f, ln)}")

# hot instructions of inlined helpers, attributed to the nearest preceding line of the kernel's own file
if len(sys.argv) > 4:
    main = sys.argv[4]
    cur = None
    byctx = defaultdict(lambda: [0, 0])
    for i in range(n):
        f, ln, txt = seq[i]
        if f == main:
            cur = ln
        if "try_wait" in txt.lower() or f == "tc_common.cuh" and ln in (28, 29, 30, 34, 35, 36):
            byctx[cur][0] += int(float(ins[i]["# Samples"] or 0)); byctx[cur][1] += int(float(ins[i]["Instructions Executed"] or 0))
    print("\nmbarrier waits by call site (nearest preceding line of %s):" % main)
    for ln, a in sorted(byctx.items(), key=lambda kv: -kv[1][0])[:30]:
        print(f"  {main}:{ln}  samples {a[0]:6d}  inst {a[1]:9d}  {text(main, ln or 1)}")
