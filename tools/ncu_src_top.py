"""print the top stall lines of an ncu --page source --csv dump.  usage: ncu_src_top.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data.sort(key=lambda r: -int(r[ix["# Samples"]] or 0))
print("total samples", tot)
for r in data[:n]:
    s = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{s:7d} {100*s/tot:5.1f}%  {r[ix['Source']][:90]:90s} {st}")
