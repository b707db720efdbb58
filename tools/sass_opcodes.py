#!/usr/bin/env python
"""Opcode mix of the library's kernels from their SASS (cuobjdump), one block per kernel: evidence of which datapath a
kernel uses (FFMA2 / HMMA / UTCHMMA + LDTM / UBLKCP / DFMA ...).  Static counts of the unrolled code, not executed counts.
    python tools/sass_opcodes.py build/obj/fwd_d1_h32.o [more .o ...] > profiles/r02/sass/<name>.txt"""
import collections, re, subprocess, sys
KEY = ["FFMA2", "FFMA", "FADD", "FMUL", "HMMA", "MOVM", "UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "MUFU", "DFMA", "DADD", "DMUL",
       "LDCU", "LDC", "LDS", "STS", "LDG", "STG", "LDL", "STL", "SHFL", "IMAD", "IADD3", "LOP3", "F2FP", "HADD2", "BAR", "ATOMG", "RED"]
for obj in sys.argv[1:]:
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); kernels[cur] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    for name, cnt in kernels.items():
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        total = sum(cnt.values())
        if total < 50:
            continue
        print(f"{obj}: {dem[:150]}")
        print(f"  instructions {total}: " + ", ".join(f"{k} {cnt[k]}" for k in KEY if cnt.get(k)))
        rest = [(k, v) for k, v in cnt.most_common() if k not in KEY][:8]
        print("  other: " + ", ".join(f"{k} {v}" for k, v in rest))
