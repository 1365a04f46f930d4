#!/usr/bin/env python
"""SASS opcode histogram per kernel of the built objects (cuobjdump -sass), so that claims such
as "DMMA.8x8x4", "LDGSTS" (cp.async), "UBLKCP" (bulk async copy) are checkable from the repo.

    tools/sass_histogram.py [build/obj/*.o ...] > profiles/rNN/sass_opcodes.txt

For every kernel: instruction count, the counts of the opcodes that carry the design
(FP64 math, tensor-core MMA, async copies, shared / global memory, barriers) and the ten most
frequent opcodes.  FP64 has no tcgen05 / UTCMMA kind, so those are expected to be absent."""
import collections
import glob
import re
import subprocess
import sys

KEY = ["DFMA", "DMUL", "DADD", "DMMA", "MUFU", "LDGSTS", "UBLKCP", "UTMALDG", "UTCMMA", "LDS",
       "STS", "LDG", "STG", "SHFL", "BAR", "WARPSYNC", "SYNCS"]


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    name, ops = None, None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            if name:
                yield name, ops
            name, ops = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Za-z0-9_]+)*)", ln)
        if m and name:
            ops[m.group(1)] += 1
            if m.group(1) in ("DMMA", "LDGSTS", "UBLKCP", "LDS", "STS", "LDG", "STG", "MUFU"):
                ops[m.group(1) + m.group(2)] += 1
    if name:
        yield name, ops


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout
    return out.splitlines()


def main():
    objs = sys.argv[1:] or sorted(glob.glob("build/obj/*.cu.o"))
    for obj in objs:
        found = list(kernels(obj))
        if not found:
            continue
        print(f"== {obj}")
        pretty = demangle([n for n, _ in found])
        for (name, ops), nice in zip(found, pretty):
            nice = re.sub(r"\(anonymous namespace\)::|sipoc::", "", nice)
            nice = re.sub(r"\(.*", "", nice)
            base = {k: v for k, v in ops.items() if "." not in k}
            total = sum(base.values())
            if total < 40:
                continue
            key = " ".join(f"{k}={ops[k]}" for k in KEY if ops.get(k))
            detail = " ".join(f"{k}={v}" for k, v in sorted(ops.items())
                              if "." in k and k.split(".")[0] in ("DMMA", "LDGSTS", "UBLKCP", "MUFU"))
            top = " ".join(f"{k}:{v}" for k, v in collections.Counter(base).most_common(10))
            print(f"{nice}\n    instructions {total} | {key}\n    {detail}\n    top: {top}")


if __name__ == "__main__":
    main()
