"""ncu launch list (`--metrics gpu__time_duration.sum --csv`) -> the compact csv (kernel, grid, block, duration_ns) and the
per-kernel markdown table kept under profiles/ (rXX_step_launches.csv / .md).
usage: python tools/launch_compact.py raw.csv out.csv out.md "<header note>" """
import collections
import csv
import sys


def main(raw, out_csv, out_md, note):
    lines = [l for l in open(raw) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    with open(out_csv, "w") as f:
        f.write(f"# {note}\nkernel,grid,block,duration_ns\n")
        w = csv.writer(f, quoting=csv.QUOTE_NONNUMERIC)
        for r in rows:
            w.writerow([r["Kernel Name"][:120], r["Grid Size"], r["Block Size"], int(float(r["Metric Value"].replace(",", "")))])
    agg = collections.defaultdict(lambda: [0, 0.0])
    lib = 0.0
    for r in rows:
        name = r["Kernel Name"].replace("void ", "")
        ms = float(r["Metric Value"].replace(",", "")) / 1e6
        key = name.split("(")[0][:90]
        agg[key][0] += 1
        agg[key][1] += ms
        if "asis::" not in name:
            lib += ms
    tot = sum(v[1] for v in agg.values())
    with open(out_md, "w") as f:
        f.write(f"{note}\n\nlaunches: {len(rows)}; sum of kernel durations: {tot:.1f} ms (ncu: serialised, cold cache -- compare "
                f"shares); library (non-asis) kernels: {lib:.2f} ms = {100 * lib / tot:.1f} %\n\n")
        f.write("| ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if v[1] / tot < 0.002:
                continue
            f.write(f"| {v[1]:.2f} | {100 * v[1] / tot:.1f}% | {v[0]} | `{k}` |\n")


if __name__ == "__main__":
    main(*sys.argv[1:5])
