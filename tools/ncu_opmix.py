"""Instruction mix + stall samples per opcode of one kernel from an .ncu-rep (source page; needs -lineinfo builds).
Usage: python tools/ncu_opmix.py report.ncu-rep kernel_regex [top]"""
import collections, csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 24
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[h]
iS, iSa, iEx = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows[h + 1:]:
    if len(r) <= iEx or not r[iEx].isdigit():
        continue
    data.append((r[iS].strip(), int(r[iSa] or 0), int(r[iEx])))
tot_ex, tot_s = sum(d[2] for d in data), sum(d[1] for d in data)
print(f"{kern}: {tot_ex} warp instructions executed, {tot_s} stall samples")
byop = collections.defaultdict(lambda: [0, 0])
for s, sa, ex in data:
    t = s.split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    byop[op][0] += ex
    byop[op][1] += sa
for op, (ex, sa) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {op:12s} executed {ex:9d} ({100 * ex / tot_ex:5.1f}%)   samples {sa:6d} ({100 * sa / tot_s:5.1f}%)")
print("  hottest instructions by samples:")
for s, sa, ex in sorted(data, key=lambda d: -d[1])[:14]:
    print(f"    {sa:6d} ({100 * sa / tot_s:4.1f}%)  x{ex:8d}  {s[:90]}")
