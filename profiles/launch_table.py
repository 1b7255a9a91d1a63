"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised launches: read the
SHARES, not the absolute times).  Usage: python profiles/launch_table.py <launches.csv>"""
import collections
import csv
import sys


def main(path):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row.get("Metric Unit", "us")
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("vgl::<unnamed>::", "").replace("unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += v
    total = sum(t for _, t in agg.values())
    print("%-44s %6s %12s %10s %6s" % ("kernel", "n", "total us", "avg us", "share"))
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %6d %12.1f %10.1f %5.1f%%" % (name[:44], n, t, t / n, 100 * t / total))


if __name__ == "__main__":
    main(sys.argv[1])
