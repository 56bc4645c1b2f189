"""Top stall locations from `ncu -i X.ncu-rep --page source --csv [--kernel-id ...]` output (a CSV file)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = next(r for r in rows if "Source" in r and "Address" in r)
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
iS, iSrc, iEx = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


tot = sum(num(r[iS]) for r in data)
print("total samples", tot, "instructions", len(data))
for r in sorted(data, key=lambda r: -num(r[iS]))[:n]:
    reasons = sorted(((num(r[i]), h[6:]) for i, h in stall_cols if num(r[i])), reverse=True)[:3]
    print(f"{num(r[iS]):7d} {100 * num(r[iS]) / max(tot, 1):5.1f}% ex={r[iEx]:>9s} {r[iSrc].strip()[:64]:64s} {reasons}")
