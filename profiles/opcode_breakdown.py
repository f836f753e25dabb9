"""Opcode-level view of one kernel from `ncu -i REPORT.ncu-rep --page source --csv` (SASS rows): share of the
executed warp instructions and of the stall samples per opcode, and the shared-memory bank-conflict wavefronts.

  python profiles/opcode_breakdown.py zeroing_source.csv[.gz] > profiles/rN_ncu_opcodes_zeroing.md
"""
import collections, csv, gzip, io, re, sys

path = sys.argv[1]
text = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
rows = list(csv.reader(io.StringIO(text)))
hdr = rows[1]
iS, iI, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iW, iWe = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive")
ops, samp, wav, wex = (collections.Counter() for _ in range(4))
for r in rows[2:]:
    if len(r) <= iWe:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS].strip())
    if not m:
        continue
    op = m.group(2)
    ops[op] += int(r[iI] or 0); samp[op] += int(r[iSm] or 0)
    wav[op] += int(r[iW] or 0); wex[op] += int(r[iWe] or 0)
tot, tots = sum(ops.values()), sum(samp.values())
cls, scl = collections.Counter(), collections.Counter()
for op, n in ops.items():
    cls[op.split(".")[0]] += n; scl[op.split(".")[0]] += samp[op]
print("# %s -- executed warp instructions by opcode (`ncu --set full --import-source on`, SASS page)\n" % rows[0][1].split("(")[0])
print("%.3e warp instructions, %d stall samples.\n" % (tot, tots))
print("| opcode | instructions | stall samples |\n|---|---:|---:|")
for b, n in cls.most_common(24):
    print("| %s | %.2f %% | %.2f %% |" % (b, 100.0 * n / tot, 100.0 * scl[b] / tots))
print("\nShared-memory wavefronts: %.3e, of which %.3e (%.0f %%) are bank-conflict replays:\n" % (sum(wav.values()), sum(wex.values()), 100.0 * sum(wex.values()) / max(1, sum(wav.values()))))
print("| opcode | wavefronts | excessive |\n|---|---:|---:|")
for op, n in wex.most_common(4):
    print("| %s | %.3e | %.3e |" % (op, wav[op], n))
