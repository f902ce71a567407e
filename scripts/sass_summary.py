#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass): the Blackwell-native evidence table of
B200_PROFILING.md -- tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA loads/stores -> UTMALDG / UTMASTG, legacy mma.sync ->
HMMA.  Usage: python scripts/sass_summary.py > profiles/r02_sass_summary.txt   (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stain2stain_b200 import _build  # noqa: E402

MNEMONICS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "HMMA", "LDSM", "MUFU", "RED", "ATOMG", "LDL", "STL"]
sass = subprocess.run(["cuobjdump", "-sass", _build.build()], capture_output=True, text=True, timeout=600).stdout
print(f"# cuobjdump -sass {os.path.relpath(_build.LIB_PATH, ROOT)}  (sm_100a)")
print(f"{'kernel':78s} {'instr':>6s} " + " ".join(f"{m:>7s}" for m in MNEMONICS))
tot = collections.Counter()
for sec in sass.split("Function : ")[1:]:
    name = sec.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
    ops = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", sec, flags=re.M)
    c = collections.Counter()
    for op in ops:
        for m in MNEMONICS:
            if op == m or (m == "HMMA" and op == "HMMA"):
                c[m] += 1
    tot.update(c)
    if any(c[m] for m in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "HMMA")):
        print(f"{dem[:78]:78s} {len(ops):6d} " + " ".join(f"{c[m]:7d}" for m in MNEMONICS))
print(f"{'TOTAL (all ' + str(len(sass.split('Function : ')) - 1) + ' kernels)':78s} {'':6s} " + " ".join(f"{tot[m]:7d}" for m in MNEMONICS))
print("# HMMA (register-level mma.sync) appears only in the attention core (csrc/attention.cuh); every conv / GEMM is UTCHMMA.")
