#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the shipped library (B200_PROFILING.md: what proves tcgen05 / TMEM /
TMA / clusters / packed FP32) -> profiles/<round>/sass_evidence.txt.

    python tools/sass_evidence.py [deeplocalproteindocking_b200/libdlpd.so] > profiles/r02/sass_evidence.txt
"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "deeplocalproteindocking_b200/libdlpd.so"
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "STTM", "LDTM", "UBLKCP", "UTMALDG", "UTMAPF", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT",
         "FFMA2", "FADD2", "FMUL2", "LDGSTS", "BAR.SYNC", "BAR.ARV", "HMMA", "ATOMG", "REDG"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
arch = re.findall(r"arch = (sm_\w+)", sass)
counts, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"^void ", "", name).split("(")[0]
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and name:
        op = m.group(1)
        counts[name]["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[name][w] += 1
print("library: %s   architectures in the fatbin: %s" % (LIB, sorted(set(arch))))
print("kernels: %d   (counts of static SASS instructions; only kernels with at least one watched mnemonic beyond barriers)\n" % len(counts))
for k, c in counts.items():
    hits = [(w, c[w]) for w in WATCH if c[w]]
    if not any(w not in ("BAR.SYNC", "BAR.ARV") for w, _ in hits):
        continue
    print("%s\n    %d instructions: %s" % (k, c["_total"], ", ".join("%s %d" % h for h in hits)))
