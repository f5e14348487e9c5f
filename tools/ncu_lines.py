#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` SASS listing per CUDA source line using nvdisasm line info.
  python tools/ncu_lines.py report.ncu-rep libzrt.so 'k_traceILi0ELi7' [--top 40]
Prints per source line: warp instructions executed, thread instructions, avg active threads, stall samples."""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os
from collections import defaultdict

rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 45
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if "kernels" in f and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
addr2line, cur, active = {}, None, False
for line in dis.splitlines():
    if line.startswith("\t.section\t.text."):
        active = kern in line
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", line)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + re.sub(r"IL[ib].*", "", kern).replace("_ZN3zrt", "")],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0, 0, 0])
base = None
tot = [0, 0, 0]
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
    if base is None:
        base = a
    key = addr2line.get(a - base, (("?", 0), ""))[0]
    ie, te, sm = int(r[col["Instructions Executed"]] or 0), int(r[col["Thread Instructions Executed"]] or 0), int(r[col["# Samples"]] or 0)
    g = agg[key]
    g[0] += ie; g[1] += te; g[2] += sm; g[3] += 1
    tot[0] += ie; tot[1] += te; tot[2] += sm
src = {}
print(f"total warp-inst {tot[0]:.4g} thread-inst {tot[1]:.4g} avg active {tot[1]/max(tot[0],1):.2f} samples {tot[2]} sass {sum(g[3] for g in agg.values())}")
for key, g in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, ln = key if key else ("?", 0)
    if f not in src and f != "?":
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "zraytrace_b200", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src.get(f, [])[ln - 1].strip()[:90] if f in src and 0 < ln <= len(src[f]) else ""
    print(f"{f}:{ln:4d} sass={g[3]:4d} winst={100*g[0]/tot[0]:5.1f}% active={g[1]/max(g[0],1):5.1f} stall={100*g[2]/max(tot[2],1):5.1f}%  {text}")
