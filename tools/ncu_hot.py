"""Top stall-sample SASS lines of an ncu report: ncu -i X.ncu-rep --page source --csv > f.csv; python tools/ncu_hot.py f.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows[:30]) if "Source" in r)
hdr = rows[h]
src, samp, inst = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
data = []
for r in rows[h + 1:]:
    if len(r) <= samp or r[0] == "Address" or r[0] == "Kernel Name":
        continue
    try:
        v = float(r[samp])
    except ValueError:
        continue
    top = sorted(((float(r[i] or 0), hdr[i][6:]) for i in stalls), reverse=True)[:2]
    data.append((v, r[src][:100], r[inst], top))
tot = sum(d[0] for d in data) or 1
print("total samples", tot)
for v, line, ni, top in sorted(data, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{v:8.0f} {100 * v / tot:5.1f}% inst={ni:>9} {top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f} | {line}")
