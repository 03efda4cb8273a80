"""Top stall sites of a kernel from `ncu --page source --csv` (SASS view): python scripts/ncu_src_top.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ia, isrc, isamp, iexe = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) > isamp and r[isamp].isdigit()]
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "instructions", len(data))
for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][isamp]))[:n]:
    why = sorted(((int(r[i]), hdr[i]) for i in stall if r[i].isdigit() and int(r[i])), reverse=True)[:2]
    print(f"{k:5d} {int(r[isamp]):6d} {100*int(r[isamp])/tot:5.1f}%  {r[isrc][:70]:70s} {why}")
