#!/usr/bin/env python
"""Markdown table of bench.py JSON lines (one file per run): python tools/bench_table.py profiles/r2_bench_n2_*.json"""
import json
import os
import sys


def load(path):
    txt = open(path).read()
    lines = [ln for ln in txt.splitlines() if ln.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def main():
    print("| run | workload | codec | N | api / schedule | transport | ms/step | GB/s | parity_ok | rel-L2 | kernels (us per launch, frac of HBM peak) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for p in sys.argv[1:]:
        d = load(p)
        if d is None:
            continue
        c = d["config"]
        ks = "; ".join(f"{k['kernel']} {k['avg_launch_us']:.1f} ({k['frac']:.2f})" for k in (d.get("roofline") or {}).get("kernels", []))
        fid = (d.get("fidelity") or {}).get("rel_l2")
        print(f"| {os.path.basename(p).replace('.json', '')} | {c['workload']} | {c['codec']} | {d['n_gpus']} | "
              f"{c.get('api', 'engine').split(' ')[0]} / {c.get('schedule', '-')} ({c.get('launch_mode', '')}) | {c.get('transport', '')} | "
              f"{d['ms_per_step']:.3f} | {d['value']:.0f} | {d.get('parity_ok')} | {fid if fid is None else round(fid, 4)} | {ks} |")


if __name__ == "__main__":
    main()
