"""Turn the raw gpurun_out/ captures into the tracked, human-readable evidence under profiles/.

    python tools/make_profiles.py r01

 - profiles/<round>_launches.csv + _launches.md: every kernel launch of the ncu launch-list pass
   with its device time, and each kernel's SHARE of one step (cold-cache, serialised: shares only)
 - profiles/<round>_<name>.ncu.txt: key metrics + top stall sites of every full capture
   gpurun_out/<round>_*.ncu-rep
 - profiles/<round>_traffic.json: dram bytes per launch of the captured kernels (bench.py's
   roofline.traffic reads this)
"""
import csv
import glob
import io
import json
import os
import re
import subprocess
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# where the summaries go: profiles/ here; on the GPU box W2E_PROFILES_OUT=gpurun_out/profiles (the .ncu-rep files are
# too big to travel back, so the summaries are made there and only the text is copied)
OUT = os.environ.get("W2E_PROFILES_OUT") or os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_hot  # noqa: E402
import ncu_summary  # noqa: E402


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name.replace("w2e::", "")[:70]


def launches(rnd):
    src = os.path.join(ROOT, "gpurun_out", f"launches_{rnd}.csv")
    if not os.path.exists(src):
        return
    rows = []
    with open(src) as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append((int(r["ID"]), r["Kernel Name"], r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e3))
    with open(os.path.join(OUT, f"{rnd}_launches.csv"), "w") as fh:
        fh.write("id,kernel,grid,block,time_us\n")
        for i, k, g, b, t in rows:
            fh.write(f'{i},"{short(k)}","{g}","{b}",{t:.2f}\n')
    # one step = the launches between two consecutive nchw_to_nhwc_mod kernels (first kernel of the engine)
    starts = [n for n, r in enumerate(rows) if "nchw_to_nhwc_mod" in r[1]]
    md = [f"# {rnd}: launch list of `python bench.py --steps 2 --warmup 3 --cpu-sample 0` under",
          "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches:",
          "compare SHARES, not absolute times)", ""]
    if len(starts) >= 2:
        seg = rows[starts[-2]:starts[-1]]
        total = sum(r[4] for r in seg)
        agg = {}
        for _, k, _, _, t in seg:
            a = agg.setdefault(short(k), [0, 0.0])
            a[0] += 1
            a[1] += t
        md += [f"One engine step = {len(seg)} launches, {total / 1e3:.2f} ms summed device time.", "",
               "| kernel | launches | time (us) | share |", "|---|---|---|---|"]
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            md.append(f"| `{k}` | {n} | {t:.1f} | {100 * t / total:.1f}% |")
    with open(os.path.join(OUT, f"{rnd}_launches.md"), "w") as fh:
        fh.write("\n".join(md) + "\n")


def full_reports(rnd):
    traffic = {}
    for rep in sorted(glob.glob(os.path.join(os.environ.get("W2E_NCU_REP_DIR") or os.path.join(ROOT, "gpurun_out"), f"{rnd}_*.ncu-rep"))):
        name = os.path.basename(rep)[:-8]
        buf = io.StringIO()
        with redirect_stdout(buf):
            print(f"# {name}: ncu --set full --clock-control none --import-source on (summary by tools/make_profiles.py)\n")
            ncu_summary.main(rep)
            print()
            ncu_hot.main(rep, 14)
        with open(os.path.join(OUT, f"{name}.ncu.txt"), "w") as fh:
            fh.write(buf.getvalue())
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr = rows[0]
        units = rows[1]

        def col(n):
            return hdr.index(n)

        def to_bytes(v, u):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            return float(v) * scale

        per = []
        for r in rows[2:]:
            rd = to_bytes(r[col("dram__bytes_read.sum")], units[col("dram__bytes_read.sum")])
            wr = to_bytes(r[col("dram__bytes_write.sum")], units[col("dram__bytes_write.sum")])
            dur = float(r[col("gpu__time_duration.sum")])
            du = units[col("gpu__time_duration.sum")]
            dur_us = dur * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(du, 1)
            per.append({"kernel": short(r[col("Kernel Name")]), "grid": r[col("Grid Size")], "dram_read_bytes": rd,
                        "dram_write_bytes": wr, "duration_us_under_ncu": dur_us})
        traffic[name] = per
    if traffic:
        with open(os.path.join(OUT, f"{rnd}_traffic.json"), "w") as fh:
            json.dump(traffic, fh, indent=1)


if __name__ == "__main__":
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    launches(rnd)
    full_reports(rnd)
