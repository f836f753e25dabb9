"""Prints (kernel, metric, value, unit) rows of an `ncu --csv` log, averaged per kernel: python ktime.py log.csv"""
import csv, sys, collections, re
agg = collections.defaultdict(lambda: [0, 0.0, ""])
for r in csv.reader(open(sys.argv[1])):
    if len(r) > 14 and r[0].isdigit():
        k = re.sub(r"\(.*", "", r[4]).replace("gzb::", "")
        try:
            v = float(r[14].replace(",", ""))
        except ValueError:
            continue
        a = agg[(k, r[12])]
        a[0] += 1; a[1] += v; a[2] = r[13]
for (k, m), (n, v, u) in sorted(agg.items()):
    print("%-34s %-28s n=%-4d avg %.3f %s" % (k, m, n, v / n, u))
