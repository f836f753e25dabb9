"""Per-source-line view of one kernel of an `ncu --set full --import-source on` capture (not a test).

  ncu -i REPORT.ncu-rep --page source --csv --kernel-name K --print-source cuda,sass > k.csv
  python profiles/source_breakdown.py k.csv [top_n]

Prints, per CUDA source line (the SASS rows under it summed): warp-stall samples, instructions executed,
shared-memory wavefronts with their ideal and excessive (bank-conflict) parts. The CSV has one section per
source file, in the order ncu lists them; the section index is printed with every line.
"""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    sec, hdr = -1, None
    agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, ""])
    line = src = None
    for r in rows:
        if not r:
            continue
        if r[0] == "Line No":
            sec, hdr = sec + 1, r
            continue
        if sec < 0 or len(r) < len(hdr):
            continue
        if r[0] != "":
            line, src = r[0], r[1]
        if r[2] == "":          # a pure source row; its SASS rows follow
            continue
        try:
            vals = [int(r[hdr.index(n)] or 0) for n in ("# Samples", "Instructions Executed", "L1 Wavefronts Shared",
                                                        "L1 Wavefronts Shared Ideal", "L1 Wavefronts Shared Excessive")]
        except ValueError:
            continue
        a = agg[(sec, line)]
        for k in range(5):
            a[k] += vals[k]
        a[5] = src.strip()
    return agg


def main():
    agg = load(sys.argv[1])
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    tot = [sum(a[k] for a in agg.values()) for k in range(5)]
    print("total: samples %d, instructions %d, shared wavefronts %d (ideal %d, excessive %d)" % tuple(tot))
    for title, key in (("by stall samples", 0), ("by excessive shared wavefronts", 4)):
        print("\n-- %s" % title)
        print("| file#:line | samples | inst | shared wf | ideal | excessive | source |")
        print("|---|---:|---:|---:|---:|---:|---|")
        for (sec, line), a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
            if a[key] == 0:
                break
            print("| %d:%s | %.1f%% | %.1f%% | %d | %d | %d | `%s` |" % (sec, line, 100.0 * a[0] / max(1, tot[0]),
                                                                 100.0 * a[1] / max(1, tot[1]), a[2], a[3], a[4], a[5][:100]))


if __name__ == "__main__":
    main()
