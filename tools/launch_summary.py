"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui], v)
        a = agg.setdefault(r[ki][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:12.1f} us {100 * t / tot:6.2f}%  n={n:5d}  avg={t / n:9.2f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
