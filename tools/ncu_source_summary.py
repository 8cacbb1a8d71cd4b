"""Condense `ncu --page source --csv` (SASS level, compiled with -lineinfo) into the part worth committing:
totals, stall-reason totals, and the instructions that collect the most samples.
    python tools/ncu_source_summary.py gpurun_out/x_source.csv > profiles/x_source_top.txt"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
kernel = rows[0][1] if rows and len(rows[0]) > 1 else "?"
hdr, data = rows[1], rows[2:]
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in data)
print(f"kernel: {kernel}")
print(f"SASS lines {len(data)}, warp-level samples {tot}, warp instructions executed {sum(int(r[iI]) for r in data)}")
agg = sorted(((hdr[i], sum(int(r[i]) for r in data)) for i in stall), key=lambda kv: -kv[1])
print("stall reasons: " + ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in agg[:8]))
print("\n  line  samples   share  executed  top stall            instruction")
top = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for i in sorted(top):
    r = data[i]
    st = sorted(((hdr[c][6:], int(r[c])) for c in stall if int(r[c]) > 0), key=lambda kv: -kv[1])
    print(f"{i:6d} {int(r[iS]):8d} {100 * int(r[iS]) / max(tot, 1):6.2f}% {int(r[iI]):9d}  {(st[0][0] if st else '-'):18s}  {r[iSrc].strip()[:90]}")
