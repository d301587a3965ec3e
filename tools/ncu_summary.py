"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total
time and share over the LAST `--steps-in-file`-th of the file (one bench step)."""
import collections
import csv
import sys


def main(path, parts=1):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = len(rows) // parts
    rows = rows[-per:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r["Kernel Name"].replace("void ", "")
        key = name.split("(")[0][:90]
        agg[key][0] += 1
        agg[key][1] += float(r["Metric Value"].replace(",", "")) / 1e6
    tot = sum(v[1] for v in agg.values())
    print(f"launches: {per}; sum of kernel durations: {tot:.1f} ms (ncu: serialised, cold cache -- compare shares)\n")
    print("| ms | share | launches | kernel |\n|---:|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.002:
            continue
        print(f"| {v[1]:.2f} | {100 * v[1] / tot:.1f}% | {v[0]} | `{k}` |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
