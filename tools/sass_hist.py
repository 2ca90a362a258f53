"""SASS instruction histogram per kernel of libvalor_b200.so (cuobjdump -sass): the mnemonics that prove which
hardware path a kernel uses (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/
UTMAREDG = TMA loads / stores / reduce-add, UBLKCP = bulk copy, HMMA = mma.sync (legacy tensor path), LDGSTS = cp.async,
SYNCS = mbarrier, MUFU = special function.   python tools/sass_hist.py > profiles/sass_histogram_r2.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "valor_b200", "csrc", "libvalor_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "LDGSTS", "SYNCS", "MUFU", "LDSM", "ATOMS", "RED", "total"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
hist = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::|valor::", "", name)
        cur = hist.setdefault(name[:110], collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        for k in KEYS:
            if op.startswith(k):
                cur[k] += 1
print(f"# {os.path.relpath(LIB, ROOT)}: SASS mnemonic counts per kernel (static instruction counts)")
print("kernel".ljust(112) + " ".join(k.rjust(8) for k in KEYS))
for name, c in hist.items():
    if c["total"] < 50:
        continue
    print(name.ljust(112) + " ".join(str(c[k]).rjust(8) for k in KEYS))
