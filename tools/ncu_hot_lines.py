#!/usr/bin/env python
"""Top source lines / SASS instructions of an `ncu --set full --import-source on` capture by warp-stall samples:
python tools/ncu_hot_lines.py REP OUT.md [N]   (run on the GPU box right after the capture: the .ncu-rep files are
too large to bring back, this text is not)."""
import csv
import io
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = []
    # the source page is a sequence of CSV tables, one per kernel launch, each starting with a header row
    blocks, cur, names, name, seen = [], [], [], "?", set()
    for ln in raw.splitlines():
        if ln.startswith('"Kernel Name"'):
            name = ln.split(",", 1)[1].strip().strip('",')[:80]
            continue
        if ln.startswith('"#"') or ln.startswith('"Source"') or ln.startswith('"Address"'):
            if cur:
                blocks.append(cur)
            cur = [ln]
            names.append(name)
        elif cur:
            cur.append(ln)
    if cur:
        blocks.append(cur)
    # one table per distinct kernel (the first launch of each)
    keep = []
    for n_, b_ in zip(names, blocks):
        if n_ not in seen:
            seen.add(n_)
            keep.append((n_, b_))
    with open(out, "w") as f:
        f.write(f"# hottest lines of `{rep}` by warp stall samples (ncu --page source)\n\n")
        for bi, (kname, blk) in enumerate(keep[:8]):
            rows = list(csv.reader(io.StringIO("\n".join(blk))))
            hdr = rows[0]
            col = {h: i for i, h in enumerate(hdr)}
            samp = next((h for h in hdr if h.startswith("# Samples") or h.startswith("Warp Stall Sampling (All")), None)
            if samp is None:
                f.write(f"(table {bi}: no sampling column; header: {hdr[:12]})\n\n")
                continue
            src = "Source" if "Source" in col else hdr[1]
            stall_cols = [h for h in hdr if h.startswith("stall_")][:24]

            def num(r, h):
                try:
                    return float(r[col[h]].replace(",", ""))
                except (ValueError, IndexError):
                    return 0.0
            data = [r for r in rows[1:] if len(r) == len(hdr)]
            tot = sum(num(r, samp) for r in data) or 1.0
            data.sort(key=lambda r: -num(r, samp))
            f.write(f"## `{kname}`: {len(data)} lines, {tot:.0f} samples\n\n| samples % | line | top stall reasons |\n|---|---|---|\n")
            for r in data[:top]:
                st = sorted(((num(r, h), h) for h in stall_cols), reverse=True)[:3]
                f.write(f"| {100 * num(r, samp) / tot:.1f} | `{r[col[src]].strip()[:110]}` | "
                        + ", ".join(f"{h.replace('stall_', '')} {v:.0f}" for v, h in st if v > 0) + " |\n")
            f.write("\n")


if __name__ == "__main__":
    main()
