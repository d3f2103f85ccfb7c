"""Turn an ncu launch-list CSV (--metrics gpu__time_duration.sum) into a per-kernel share table (markdown)."""
import collections
import csv
import sys


def table(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        agg[r[ki].split("(")[0].replace("void ", "")[:70]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out = ["| kernel | launches | avg us | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {sum(v) / tot:.3f} |")
    return "\n".join(out)


if __name__ == "__main__":
    print(table(sys.argv[1]))
