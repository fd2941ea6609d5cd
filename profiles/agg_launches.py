"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = row["Metric Unit"]
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:110]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"total {tot / 1e3:.1f} us over {sum(c for c, _ in agg.values())} launches")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{t / 1e3:10.1f} us {100 * t / tot:5.1f}%  n={c:5d}  avg={t / c / 1e3:8.2f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
