"""Per-kernel SASS opcode histogram of csrc/libgwsim.so (cuobjdump -sass), written to profiles/.  Shows at a glance which kernels
use TMA bulk copies (UBLKCP), warp reductions (REDUX), 128-bit loads / stores, fp64 (DMUL / DADD / DFMA) and that nothing here
touches the tensor cores (no HMMA / UTCMMA: nothing on this path is a contraction).

    python scripts/sass_histogram.py > profiles/r02_sass_histograms.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ai_safety_gridworlds_b200", "csrc", "libgwsim.so")
text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kernels = collections.OrderedDict()
name = None
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        kernels[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
    if m and name:
        kernels[name][m.group(1)] += 1
INTEREST = ("UBLKCP", "REDUX", "LDG.E.128", "STG.E.128", "LDS.128", "STS.128", "ATOMS", "ATOMG", "RED", "DMUL", "DADD", "DFMA", "MUFU", "HMMA", "UTCMMA", "SHFL", "VOTE")
print("SASS of %s (sm_100a), instructions per kernel and the opcodes that matter here" % os.path.relpath(lib, ROOT))
for k, c in kernels.items():
    total = sum(c.values())
    fam = collections.Counter()
    for op, n in c.items():
        for key in INTEREST:
            if op.startswith(key) or (key in ("LDG.E.128", "STG.E.128") and op.startswith(key.split(".")[0]) and ".128" in op):
                fam[key] += n
                break
    print("\n%s\n  %d instructions (%.1f KB); %s" % (k[:150], total, total * 16 / 1024.0, ", ".join("%s %d" % kv for kv in sorted(fam.items()))))
    print("  top opcodes: " + ", ".join("%s %d" % kv for kv in c.most_common(14)))
