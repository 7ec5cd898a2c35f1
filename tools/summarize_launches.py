"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step).

    python tools/summarize_launches.py gpurun_out/launches.csv [> profiles/rNN_launches.md]
"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, data = rows[0], rows[1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in data:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        d = agg.setdefault(r[ki], [0, 0.0])
        d[0] += 1
        d[1] += v
        tot += v
    print("| share | total us | launches | avg us | kernel |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| %5.2f%% | %9.1f | %5d | %7.1f | `%s` |" % (t / tot * 100, t, n, t / n, k[:120]))
    print("\ntotal %.1f us over %d launches (ncu per-launch times are cold-cache and serialised: compare shares)" % (tot, len(data)))


if __name__ == "__main__":
    main(sys.argv[1])
