#!/usr/bin/env python
"""Summarise ncu captures into the small tracked files under profiles/.

  python tools/ncu_summary.py full  gpurun_out/X.ncu-rep  profiles/NAME.md  [--traffic KEY_SUFFIX]
      per-launch key metrics of an `ncu --set full` capture (+ DRAM bytes into profiles/traffic.json)
  python tools/ncu_summary.py list  gpurun_out/launches.csv profiles/NAME.md
      per-kernel time shares of an `ncu --metrics gpu__time_duration.sum` launch list
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of ncu peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("smsp__inst_executed.sum", "warp insts"),
]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def full(rep, out_md, traffic_suffix=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines = ["| # | kernel | " + " | ".join(n for _, n in METRICS) + " | top stalls (warps per issue) |",
             "|---|---|" + "---|" * (len(METRICS) + 1)]
    traffic = {}
    for n, r in enumerate(rows[2:]):
        name = r[col["Kernel Name"]].replace("void ", "").split("(")[0]
        cells = []
        for m, _ in METRICS:
            if m in col:
                v, u = r[col[m]], units[col[m]]
                try:
                    cells.append(f"{float(v.replace(',', '')):.4g} {u}".strip())
                except ValueError:
                    cells.append(v)
            else:
                cells.append("-")
        stalls = []
        for h, i in col.items():
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        top = ", ".join(f"{h} {v:.2f}" for v, h in sorted(stalls, reverse=True)[:4])
        lines.append(f"| {n} | `{name}` | " + " | ".join(cells) + f" | {top} |")
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        base = name.split("<")[0].replace("cf::", "")
        traffic.setdefault(base, []).append(rd + wr)
    with open(out_md, "w") as f:
        f.write(f"# ncu --set full summary of `{os.path.basename(rep)}`\n\n"
                "Cold-cache, serialised replays (`--clock-control none`): durations here are for share / counter\n"
                "analysis, not bench values.\n\n" + "\n".join(lines) + "\n")
    if traffic_suffix:
        path = os.path.join(ROOT, "profiles", "traffic.json")
        cur = json.load(open(path)) if os.path.exists(path) else {}
        for k, v in traffic.items():
            cur[f"{k}|{traffic_suffix}"] = sum(v) / len(v)
        json.dump(cur, open(path, "w"), indent=1, sort_keys=True)


def launch_list(csv_path, out_md):
    with open(csv_path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        name = row["Kernel Name"].replace("void ", "").split("(")[0]
        agg[(name, row["Grid Size"], row["Block Size"])].append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out_md, "w") as f:
        f.write(f"# ncu launch list `{os.path.basename(csv_path)}` (gpu__time_duration.sum, --clock-control none)\n\n"
                "Per-launch times are cold-cache and serialised: compare SHARES with bench.py, not absolutes.\n\n"
                "| kernel | grid | block | launches | avg us | min us | share |\n|---|---|---|---|---|---|---|\n")
        for (name, grid, block), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{name}` | {grid} | {block} | {len(v)} | {sum(v) / len(v):.2f} | {min(v):.2f} | {sum(v) / tot * 100:.1f}% |\n")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "full":
        suffix = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
        full(sys.argv[2], sys.argv[3], suffix)
    else:
        launch_list(sys.argv[2], sys.argv[3])
