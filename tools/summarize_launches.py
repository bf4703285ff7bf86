"""ncu launch list (gpu__time_duration.sum, --csv) -> per-kernel share table (markdown)."""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    n = 0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e6 if unit == "ns" else v / 1e3 if unit == "us" else v
        key = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("rvb::", "")[:60]
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    print(f"| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {v:.3f} | {100 * v / tot:.1f} % |")
    print(f"| **all** | {n} | {tot:.3f} | 100 % |")


if __name__ == "__main__":
    main(sys.argv[1])
