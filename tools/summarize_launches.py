"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean duration, share."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in data:
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        agg[r[ki][:72]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':72s} {'n':>4s} {'mean us':>9s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:72s} {len(v):4d} {sum(v) / len(v):9.1f} {sum(v) / tot * 100:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
