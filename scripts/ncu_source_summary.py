"""Opcode / region histogram of an `ncu --page source --csv --print-source sass` export: where the issue slots and the
stall samples of a kernel go.  usage: ncu_source_summary.py file.csv [n_top_instructions]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = next(r for r in rows if r and r[0] == "Address")
body = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
iI, iS, iSrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
tot = sum(int(r[iI]) for r in body)
ts = max(1, sum(int(r[iS]) for r in body))
print(f"{tot} warp instructions, {ts} samples, {len(body)} SASS lines")
ops, smp = collections.Counter(), collections.Counter()
for r in body:
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[iSrc])
    op = m.group(2).split(".")[0] if m else r[iSrc][:10]
    ops[op] += int(r[iI])
    smp[op] += int(r[iS])
for op, c in ops.most_common(24):
    print(f"  {op:10s} {c:11d} {100 * c / tot:5.1f} %   samples {100 * smp[op] / ts:5.1f} %")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.Counter()
for r in body:
    for i in stall:
        agg[hdr[i]] += int(r[i] or 0)
print("  stalls:", ", ".join(f"{k[6:]} {100 * v / ts:.0f}%" for k, v in agg.most_common(8)))
if top:
    for r in sorted(body, key=lambda r: -int(r[iS]))[:top]:
        print(f"  {int(r[iS]):6d} smp {int(r[iI]):9d} x  {r[0][-5:]} {r[iSrc].strip()[:90]}")
