"""Top stall locations of each kernel in an .ncu-rep (source page, SASS view): prints the
instructions with the most warp-stall samples and their dominant stall reason."""
import csv
import subprocess
import sys


def main(path, topn=22, kernel_regex=None):
    cmd = ["ncu", "-i", path, "--page", "source", "--csv"]
    if kernel_regex:
        cmd += ["--kernel-name", "regex:" + kernel_regex]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(out.splitlines()):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None and row:
            cur["rows"].append(row)
    for k, blk in enumerate(blocks):
        hdr = blk["hdr"]
        si = hdr.index("# Samples")
        src = hdr.index("Source")
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        total = sum(int(r[si] or 0) for r in blk["rows"])
        print(f"=== launch {k}: {blk['name'][:60]}  total samples {total}")
        agg = {}
        for r in blk["rows"]:
            for i in stall_cols:
                agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
        print("   stall mix:", ", ".join(f"{n[6:]} {100 * v / max(total, 1):.0f}%" for n, v in sorted(agg.items(), key=lambda x: -x[1])[:7]))
        rows = sorted(blk["rows"], key=lambda r: -int(r[si] or 0))[:topn]
        for r in rows:
            n = int(r[si] or 0)
            top = max(stall_cols, key=lambda i: int(r[i] or 0))
            print(f"   {100 * n / max(total, 1):5.1f}%  {hdr[top][6:]:12s} {r[src][:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 22, sys.argv[3] if len(sys.argv) > 3 else None)
