"""SASS opcode histogram of every kernel in libtbknarpe.so (runs in the build container, no GPU):
    python profiles/sass_opcodes.py > profiles/r2/sass_opcodes.txt
Proof of what the kernels are made of: UTCHMMA / UTCQMMA (tcgen05.mma), UTMALDG (TMA loads), LDTM (tcgen05.ld),
UTCBAR (tcgen05.commit), HMMA.16816 (legacy mma.sync), MOVM (movmatrix), MUFU, SYNCS (mbarrier) ..."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "trafficbotsv1.5_b200", "libtbknarpe.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").replace("void ", "")
        depth, cut = 0, len(kern)
        for i, ch in enumerate(kern):  # cut the argument list: the first "(" outside template brackets
            if ch == "<":
                depth += 1
            elif ch == ">":
                depth -= 1
            elif ch == "(" and depth == 0:
                cut = i
                break
        kern = kern[:cut]
        hist[kern] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        op = m.group(1)
        key = op.split(".")[0]
        if key in ("HMMA", "UTCHMMA", "UTMALDG", "LDTM", "LDG", "STG", "MUFU", "SYNCS", "UTCBAR", "FENCE", "LDGSTS", "UBLKCP"):
            key = ".".join(op.split(".")[:3]) if key in ("HMMA", "LDG", "LDTM") else ".".join(op.split(".")[:2])
        hist[kern][key] += 1
print("architectures:", ", ".join(sorted(arch)))
KEY = ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "HMMA", "MOVM", "MUFU", "SYNCS", "SHFL", "REDUX", "LDG", "STG", "LDS", "STS", "FFMA", "F2FP")
for k, c in hist.items():
    tot = sum(c.values())
    sel = {o: n for o, n in c.items() if o.split(".")[0] in KEY}
    print(f"\n{k}  ({tot} instructions)")
    print("   " + "  ".join(f"{o}:{n}" for o, n in sorted(sel.items(), key=lambda kv: -kv[1])))
