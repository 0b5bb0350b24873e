"""SASS mnemonic histogram of libsegp.so per kernel (cuobjdump -sass): the evidence that the shipped binary holds
tcgen05 MMAs (UTCIMMA), TMEM loads (LDTM), bulk copies (UBLKCP), mbarrier transactions (SYNCS), FP64 tensor
instructions (DMMA).  Writes profiles/round2/sass_histogram.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "safe_exploration_b200", "libsegp.so")
OUT = os.path.join(ROOT, "profiles", "round2", "sass_histogram.txt")
WATCH = ("UTCIMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "DMMA", "DFMA", "DADD", "DMUL", "PRMT", "SHFL", "LDG", "STG", "LDS",
         "STS", "BAR", "UCGABAR", "F2I", "I2F")

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
if "Function :" not in out:
    sys.exit("cuobjdump produced no SASS")
per = collections.OrderedDict()
cur = None
arch = set(re.findall(r"arch = (sm_\w+)", out))
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
    if m and cur is not None:
        op = m.group(1)
        per[cur]["__total__"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                per[cur][w] += 1
        if op.startswith("UTCIMMA") and ".2CTA" in op:
            per[cur]["UTCIMMA.2CTA"] += 1
        if op.startswith("UBLKCP") and "MULTICAST" in op:
            per[cur]["UBLKCP.MULTICAST"] += 1
with open(OUT, "w") as f:
    f.write("# cuobjdump -sass safe_exploration_b200/libsegp.so  (arch: %s); instructions per kernel, selected mnemonics\n"
            % ", ".join(sorted(arch)))
    cols = list(WATCH) + ["UBLKCP.MULTICAST"]
    f.write("kernel, total, " + ", ".join(cols) + "\n")
    for k, c in per.items():
        if c["__total__"] == 0:
            continue
        f.write("%s, %d, %s\n" % (k, c["__total__"], ", ".join(str(c[w]) for w in cols)))
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    f.write("ALL, %d, %s\n" % (tot["__total__"], ", ".join(str(tot[w]) for w in cols)))
print(open(OUT).read()[:3000])
