"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, gi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        agg.setdefault((r[ki][:64], r[gi]), []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':64s} {'grid':>16s} {'n':>4s} {'avg_us':>9s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k[0]:64s} {k[1]:>16s} {len(v):4d} {sum(v) / len(v) / 1e3:9.1f} {100 * sum(v) / tot:6.1f}%")
    print(f"total {tot / 1e3:.1f} us over {sum(len(v) for v in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1])
