"""Summarises an ncu launch list (gpu__time_duration.sum CSV) per kernel: launches, total and share."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        a = agg.setdefault(r[ik].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values())
    print("%-72s %6s %12s %7s" % ("kernel", "n", "total ms", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %6d %12.3f %6.1f%%" % (k, n, t, 100 * t / tot))
    print("%-72s %6s %12.3f" % ("all", "", tot))


if __name__ == "__main__":
    main(sys.argv[1])
