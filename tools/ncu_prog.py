"""Per-instruction execution counts and stall samples of one kernel launch of an .ncu-rep (source
page): prints the instruction-count classes (how often each loop level runs), the opcode mix of a
chosen class and its hottest stall sites.   python tools/ncu_prog.py REP LAUNCH_INDEX [COUNT]"""
import csv
import re
import subprocess
import sys
from collections import Counter


def main(path, launch, count=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    idx = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    idx.append(len(rows))
    print("kernel:", rows[idx[launch]][1][:120])
    hdr = rows[idx[launch] + 1]
    blk = rows[idx[launch] + 2:idx[launch + 1]]
    ie, si, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    prog = [(i, int(r[ie] or 0), int(r[si] or 0), r[src]) for i, r in enumerate(blk) if len(r) > src]
    tot = sum(p[1] for p in prog)
    ts = sum(p[2] for p in prog)
    print(f"executed {tot}  samples {ts}")
    cls = Counter()
    smp = Counter()
    for _, n, s, _ in prog:
        cls[n] += 1
        smp[n] += s
    for n, k in sorted(cls.items(), key=lambda x: -x[0] * x[1])[:8]:
        print(f"  count {n:10d} x {k:4d} instrs = {100 * n * k / tot:5.1f}% of executed, {100 * smp[n] / max(ts, 1):5.1f}% of samples")
    if count is None:
        count = max(cls.items(), key=lambda x: x[0] * x[1])[0]
    sel = [p for p in prog if p[1] == count]
    ops = Counter(re.sub(r"^@!?U?P\d+\s+", "", p[3].strip()).split()[0].split(".")[0] for p in sel)
    print(f"class {count}: {len(sel)} instrs:", ops.most_common(30))
    print("hottest:")
    for i, n, s, t in sorted(prog, key=lambda p: -p[2])[:25]:
        print(f"  {i:5d} {n:9d} {100 * s / max(ts, 1):5.1f}% {t[:100]}")
    return prog


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else None)
