"""Write a markdown summary (key metrics per kernel) of one or more .ncu-rep files.
    python tools/ncu_md.py "title" out.md rep1.ncu-rep [rep2 ...]"""
import csv
import subprocess
import sys

from ncu_key import WANT


def main(title, out, reps):
    lines = [f"# {title}", ""]
    for path in reps:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        lines += [f"## `{path.split('/')[-1]}`", ""]
        for r in rows[2:]:
            lines += [f"### `{r[hdr.index('Kernel Name')][:100]}`  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}", ""]
            for w in WANT:
                if w in hdr and r[hdr.index(w)] not in ("", "n/a"):
                    i = hdr.index(w)
                    lines.append(f"- {w}: {r[i]} {units[i]}")
            lines.append("")
    open(out, "w").write("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3:])
