#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove the tcgen05 / TMEM / TMA path (B200_PROFILING.md):
UTCHMMA (tcgen05.mma, .2CTA = cta_group::2), LDTM (tcgen05.ld), UTMALDG / UTMASTG / UTMAPF (TMA load / store /
prefetch), UTCBAR (tcgen05.commit), SYNCS (mbarrier).  Usage: python tools/sass_opcodes.py [lib.so] > profiles/r02_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "facerecognitionpipeline_b200/libfrb200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
pat = re.compile(r"\b(UTC[A-Z]*MMA[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UTMAPF[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|SYNCS[.\w]*|HMMA[.\w]*)")
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = per.setdefault(m.group(1), collections.Counter())
        continue
    if cur is not None:
        for op in pat.findall(line):
            cur[op] += 1
print(f"# SASS opcode evidence for {lib} (cuobjdump -sass, sm_100a); kernels without any of the opcodes are listed last")
plain = []
for fn, c in per.items():
    full, depth, cut = demangle(fn), 0, None
    for i, ch in enumerate(full):          # the parameter list starts at the first "(" outside template brackets
        depth += ch == "<"
        depth -= ch == ">"
        if ch == "(" and depth == 0 and not full[:i].endswith("<unnamed>") and "unnamed" not in full[max(0, i - 1):i + 9]:
            cut = i
            break
    name = (full[:cut] if cut else full).replace("void frb::", "").replace("frb::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    if not c:
        plain.append(name)
        continue
    print(f"\n## {name}")
    for op, n in sorted(c.items()):
        print(f"{n:7d} {op}")
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print("\n## whole library")
for op, n in sorted(tot.items()):
    print(f"{n:7d} {op}")
print("\n## kernels with none of these opcodes (CUDA-core kernels): " + ", ".join(sorted(set(plain))))
