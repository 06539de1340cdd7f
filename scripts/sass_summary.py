#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show the Blackwell-native paths (B200_PROFILING.md "What proves a Blackwell-native
kernel") in the shipped library: writes profiles/<tag>_sass.txt.   python scripts/sass_summary.py r02"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(ROOT, "sdface-gan_b200", "lib", "libsdfg.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "MUFU.SIN", "MUFU.SQRT", "FFMA2", "FMUL2", "RED.E", "HMMA"]
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0][:110]
        counts[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    for p in pats:
        if re.search(r"\b" + re.escape(p) + r"\b", line):
            counts[cur][p] += 1
            if p == "UTCHMMA.2CTA":
                break
out = ["# SASS mnemonic counts per kernel of sdface-gan_b200/lib/libsdfg.so (cuobjdump -sass), kernels with tcgen05 / TMA / packed-fp32 code first",
       "%-112s " % "kernel" + " ".join("%12s" % p for p in pats)]
rows = sorted(counts.items(), key=lambda kv: -(kv[1]["UTCHMMA"] + kv[1]["UTCHMMA.2CTA"] + kv[1]["UTMALDG"]))
for k, c in rows:
    if sum(c.values()) == 0:
        continue
    out.append("%-112s " % k + " ".join("%12d" % c[p] for p in pats))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
out.append("%-112s " % "TOTAL" + " ".join("%12d" % tot[p] for p in pats))
dst = os.path.join(ROOT, "profiles", "%s_sass.txt" % tag)
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))
