"""Counts of the Blackwell-native SASS instructions per kernel of libbla.so (no GPU needed: `cuobjdump -sass`).
Usage: python profiles/sass_summary.py > profiles/rNN_sass_gemm_tc.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "big-linear-algebra_b200", "libbla.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
WANT = ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "STAS", "ELECT", "UTMACCTL", "ACQBULK", "PREEXIT", "UCGABAR")
per = collections.OrderedDict()
cur = None
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for w in WANT:
            if op.startswith(w):
                key = "UTCHMMA.2CTA" if (w == "UTCHMMA" and ".2CTA" in op) else w
                per[cur][key] += 1
print("cuobjdump -sass big-linear-algebra_b200/libbla.so (sm_100a), Blackwell-native instruction counts per kernel")
print("mnemonics: UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store,")
print("UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, STAS = st.async, ELECT = elect.sync, UCGABAR = barrier.cluster,")
print("ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents (programmatic dependent launch)\n")
total = collections.Counter()
for k, c in per.items():
    tc = {w: n for w, n in c.items() if w not in ("ACQBULK", "PREEXIT")}
    if not tc and not c:
        continue
    if any(w in c for w in ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UBLKCP", "STAS", "UCGABAR")) or c.get("ACQBULK"):
        print(k[:170])
        print("    " + ", ".join(f"{w} {n}" for w, n in c.items()))
        total.update(c)
print("\ntotal: " + ", ".join(f"{w} {n}" for w, n in total.items()))
