#!/usr/bin/env python
"""Static SASS instruction count per CUDA source line of one kernel (nvdisasm --print-line-info on the built object).
  python tools/sass_lines.py zraytrace_b200/csrc/build/zrt_kernels.o 'k_trace_bpoolILi128' [--min 3]
Inlined helpers are attributed to their own source lines.  Used with the emulator's section counts (tools/emu) to
estimate warp instructions per ray of a scheduler variant before spending GPU time."""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

obj, kern = sys.argv[1], sys.argv[2]
mn = int(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 1
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cnt, cur, active, total = Counter(), None, False, 0
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
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
        cnt[cur] += 1
        total += 1
print(f"# {kern}: {total} SASS instructions")
for (f, l), c in sorted(cnt.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if c >= mn:
        print(f"{f}:{l}\t{c}")
