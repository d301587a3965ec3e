"""Top SASS instructions by warp-stall samples from an .ncu-rep (source page), with neighbours."""
import csv
import subprocess
import sys


def main(path, kernel_filter=None, top=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(out.splitlines()):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": [], "hdr": None}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None:
            cur["rows"].append(row)
    for b in blocks:
        if kernel_filter and kernel_filter not in b["name"]:
            continue
        hdr = b["hdr"]
        si, src = hdr.index("# Samples"), hdr.index("Source")
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(float(r[si] or 0) for r in b["rows"] if len(r) > si)
        print("==", b["name"][:100], "total samples", tot)
        idx = sorted(range(len(b["rows"])), key=lambda i: -float(b["rows"][i][si] or 0) if len(b["rows"][i]) > si else 0)[:top]
        for i in sorted(idx):
            r = b["rows"][i]
            v = float(r[si] or 0)
            st = sorted(((float(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
            print(f"  {i:5d} {v:7.0f} {100 * v / tot:5.1f}%  {r[src][:70]:70s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, int(sys.argv[3]) if len(sys.argv) > 3 else 25)
