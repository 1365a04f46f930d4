#!/usr/bin/env python
"""Per-source-line stall table of one kernel from an ncu capture.

    tools/stalls_by_line.py <report.ncu-rep> <object.o|.cubin> <mangled-name-substring> [top]

ncu's `--page source --csv` lists the kernel's SASS in address order with the stall samples
per instruction; `nvdisasm -g` lists the same instructions with `//## File ... line N`
markers (compile with -lineinfo).  The two are matched by position.
"""
import csv
import io
import re
import subprocess
import sys
import collections
import os
import tempfile

rep, obj, sub = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
if obj.endswith(".cubin"):
    cubin = obj
else:
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp,
                          stdout=subprocess.DEVNULL)
    cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines, cur, inside, inl = [], None, False, None
src_file = None
for ln in sass.splitlines():
    if ln.startswith(".text."):
        inside = sub in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1), int(m.group(2)))
        if src_file is None or m.group(1).endswith((".cu", ".cuh")):
            src_file = m.group(1) if m.group(1).endswith((".cu", ".cuh")) else src_file
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
# NCU_KERNEL=<regex> picks one kernel of a report that holds several
pick = ["-k", "regex:" + os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + pick, capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = rows[hdr_i + 1:]
print(f"sass rows {len(body)} disasm instrs {len(lines)}")
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.defaultdict(lambda: collections.Counter())
n = min(len(body), len(lines))
for r, ln in zip(body[:n], lines[:n]):
    a = agg[ln]
    a["samples"] += float(r[col["# Samples"]] or 0)
    a["inst"] += float(r[col["Instructions Executed"]] or 0)
    for s in stalls:
        a[s] += float(r[col[s]] or 0)
tot = sum(a["samples"] for a in agg.values()); toti = sum(a["inst"] for a in agg.values())
print(f"total samples {tot} inst {toti}")
srcs = {}
def text_of(key):
    if not key:
        return ""
    f, ln = key
    if f not in srcs:
        srcs[f] = open(f).read().splitlines() if os.path.exists(f) else []
    t = srcs[f][ln - 1].strip()[:90] if ln <= len(srcs[f]) else ""
    return t if f.endswith((".cu", ".cuh")) else f"[{os.path.basename(f)}] {t}"
keys = ["stall_wait", "stall_short_sb", "stall_long_sb", "stall_math", "stall_selected",
        "stall_mio", "stall_no_inst", "stall_barrier", "stall_lg", "stall_not_selected"]
print("line samp% inst% " + " ".join(k[6:11].rjust(6) for k in keys) + " | src")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    text = text_of(ln)
    print(f"{(ln[1] if ln else 0)!s:>5} {100*a['samples']/tot:5.1f} {100*a['inst']/toti:5.1f} " +
          " ".join(f"{a[k]:6.0f}" for k in keys) + " | " + text)
# coarse split by line range if given
if len(sys.argv) > 5:
    lo, hi = map(int, sys.argv[5].split("-"))
    ok = lambda ln: ln and ln[0].endswith(".cu") and lo <= ln[1] <= hi
    s = sum(a["samples"] for ln, a in agg.items() if ok(ln))
    i = sum(a["inst"] for ln, a in agg.items() if ok(ln))
    print(f"lines {lo}-{hi}: {100*s/tot:.1f}% of samples, {100*i/toti:.1f}% of instructions")
