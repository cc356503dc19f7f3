"""Aggregate an `ncu --page source --csv` dump: instruction mix and where the samples fall.
usage: ncu -i X.ncu-rep --page source --csv | python scripts/ncu_sass_hot.py [block]"""
import collections
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr = next(r for r in rows if r and r[0] == "Address")
data, seen = [], set()
for r in rows:
    if r and r[0].startswith("0x") and len(r) > 10:
        if r[0] in seen:      # further kernel instances repeat the addresses: keep the first
            break
        seen.add(r[0])
        data.append(r)
ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
tot = sum(int(r[ia]) for r in data)
stot = sum(int(r[isamp]) for r in data)
print("SASS instructions", len(data), "executed (warp-level)", tot, "samples", stot)
ops, samp = collections.Counter(), collections.Counter()
for r in data:
    t = r[isrc].split()
    op = t[1] if t[0].startswith("@") else t[0]
    ops[op] += int(r[ia])
    samp[op] += int(r[isamp])
for op, c in ops.most_common(30):
    print(f"{op:24s} {c:11d} {100 * c / tot:5.1f}%   samples {samp[op]:7d} {100 * samp[op] / max(stot, 1):5.1f}%")
blk = int(sys.argv[1]) if len(sys.argv) > 1 else 250
print("--- by SASS position ---")
for i in range(0, len(data), blk):
    c = sum(int(r[ia]) for r in data[i:i + blk])
    s = sum(int(r[isamp]) for r in data[i:i + blk])
    print(f"{i:6d} exec {100 * c / tot:5.1f}%  samples {100 * s / max(stot, 1):5.1f}%   {data[i][isrc][:50]}")
