"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples, and samples grouped by opcode."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iexe = hdr.index('Address'), hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
data = []
for r in rows[2:]:
    try:
        data.append((int(r[isamp]), int(r[iexe]), r[isrc].strip()))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
byop = collections.Counter(); exe = collections.Counter()
for s, e, src in data:
    op = src.split()[0] if not src.startswith('@') else src.split()[1]
    op = op.split('.')[0]
    byop[op] += s; exe[op] += e
print("--- samples by opcode (share of samples, executed warp-instructions)")
for op, s in byop.most_common(14):
    print("%-10s %6.2f%%  exec=%d" % (op, 100.0 * s / tot, exe[op]))
print("--- top instructions")
for s, e, src in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%6.2f%%  exec=%-9d %s" % (100.0 * s / tot, e, src))
