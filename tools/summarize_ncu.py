"""Turns an `ncu --set full` capture into profiles/<tag>_ncu_summary.md and profiles/traffic.json.
Input: the .ncu-rep, or the `ncu -i <rep> --page raw --csv` export of it (the full report of the
whole suite exceeds gpurun's 64 MiB return limit, so the export is made on the GPU box).
Usage: python tools/summarize_ncu.py gpurun_out/prof_r1_raw.csv r1"""
import csv
import io
import json
import os
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if rep.endswith(".csv"):
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return float("nan")


def scale(name):  # ncu prints adaptive units
    u = units[col[name]]
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)


lines = ["# ncu --set full summary (" + tag + ")", "",
         "Source: `" + rep + "` (captured with `--clock-control none --import-source on`, one B200; kernel replay,"
         " so durations are cold-cache and serialised: compare shares and DRAM bytes, not absolutes).", "",
         "| # | kernel | grid | block | duration ms | DRAM read GB | DRAM write GB | DRAM GB/s (this replay) | dram % of ncu peak | achieved occupancy % | regs | issue slots busy % |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|"]
traffic = {}
for i, r in enumerate(data):
    name = r[col["Kernel Name"]].replace("void ", "").split("(")[0]
    dur = val(r, "gpu__time_duration.sum") * scale("gpu__time_duration.sum")
    rd = val(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum")
    wr = val(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum")
    lines.append(f"| {i} | `{name}` | {r[col['Grid Size']]} | {r[col['Block Size']]} | {dur * 1e3:.4f} | {rd / 1e9:.4f} | {wr / 1e9:.4f} | "
                 f"{(rd + wr) / dur / 1e9:.0f} | {val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                 f"{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {int(val(r, 'launch__registers_per_thread'))} | "
                 f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} |")
    traffic.setdefault(name, []).append(rd + wr)
out_md = os.path.join(root, "profiles", f"{tag}_ncu_summary.md")
open(out_md, "w").write("\n".join(lines) + "\n")
# per-launch DRAM traffic of the dominant kernel (bench.py's roofline.traffic): average over its launches
dom = [v for k, vs in traffic.items() if "ew_nd_kernel<4, 4" in k and "BinF<0>" in k for v in vs]
# bench.py averages ONE c2_add launch and ONE c3_bcast_add_vec launch; a capture may hold several of either
small, big = [v for v in dom if v < 10e9], [v for v in dom if v >= 10e9]
dom_avg = (sum(small) / len(small) + sum(big) / len(big)) / 2 if small and big else (sum(dom) / len(dom) if dom else None)
tj = {"source": os.path.basename(rep), "ew_nd_vec4_add_bytes_per_launch": int(dom_avg) if dom_avg else None,
      "per_kernel_bytes": {k: [int(x) for x in v] for k, v in traffic.items()}}
json.dump(tj, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
print(open(out_md).read())
