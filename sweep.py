#!/usr/bin/env python
"""sweep.py -- kernel sweep (BASELINE.json configs[4]): every codec of the hot path over
1 MB - 1 GB fp16 activations, achieved algorithmic GB/s against the measured HBM roofline.

Each point times ONE C-ABI call (all the launches that call makes) on device-resident
inputs.  Cold-cache discipline: the call is issued over R rotating buffer sets whose total
footprint exceeds 2x the 126 MB L2, the R calls are captured in one CUDA graph (no CPU
launch gaps), and the graph is replayed `--reps` times between two CUDA events.

Algorithmic bytes per call are SURVEY.md section 8(d)'s figures (E = N*C fp16 elements):
  compress+EF : 2E (x) + 2E (base) + 2E (new_base) + code + scales
  compress    : 2E (x) + 2E (base) + code + scales            (all-gather sender)
  decompress  : 2E (base) + code + scales + 2E (recon)

  python sweep.py [--sizes-mb 1,8,27,64,256,1024] [--ops binary,int2,int4,int8,topk,lowrank]
                  [--out gpurun_out/sweep.jsonl] [--md profiles/sweep.md]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from compactfusion_b200 import _native as nv  # noqa: E402

L2_BYTES = 126 << 20
C_FIXED = 3072


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class Point:
    """One (op, N, C) measurement: `make(i)` returns a zero-arg launcher over buffer set i."""

    def __init__(self, name, n, c, algo_bytes, footprint, make, payload_bytes=None):
        self.name, self.n, self.c = name, n, c
        self.algo_bytes, self.footprint, self.make, self.payload_bytes = algo_bytes, footprint, make, payload_bytes


def _rand_pair(n, c, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(n, c, generator=g, device=dev)
    base = (x + 0.3 * torch.randn(n, c, generator=g, device=dev)).half()
    return x.half(), base


def sign_points(codec_name, n, c, dev):
    codec = nv.CODEC_BINARY if codec_name == "binary" else nv.CODEC_INT2
    per_byte = 8 if codec_name == "binary" else 4
    e = n * c
    code, scales = e // per_byte, 2 * (n + c)
    lib = nv.lib()
    cfn = lib.cf_binary_compress if codec_name == "binary" else lib.cf_int2_compress
    ws_bytes = nv.workspace_bytes(codec, n, c)

    def mk_compress(update):
        def make(i):
            x, base = _rand_pair(n, c, dev, i)
            packed = torch.empty(code, dtype=torch.uint8, device=dev)
            u = torch.empty(n, dtype=torch.half, device=dev)
            v = torch.empty(c, dtype=torch.half, device=dev)
            nb = torch.empty_like(x) if update else None
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            keep = (x, base, packed, u, v, nb, ws)

            def run():
                nv.check(cfn(nv.ptr(x), nv.ptr(base), nv.ptr(nb), nv.ptr(packed), nv.ptr(u), nv.ptr(v), n, c,
                             nv.ptr(ws), ws.numel(), nv.stream_ptr()), "compress")
            run.keep = keep
            return run
        return make

    def make_dec(i):
        x, base = _rand_pair(n, c, dev, i)
        packed = torch.randint(0, 256, (code,), dtype=torch.uint8, device=dev)
        u = torch.rand(n, device=dev).half()
        v = torch.rand(c, device=dev).half()
        recon = torch.empty_like(x)

        def run():
            if codec_name == "binary":
                rc = lib.cf_binary_decompress(nv.ptr(packed), nv.ptr(u), nv.ptr(v), 1, nv.ptr(base), nv.ptr(recon), n, c,
                                              nv.stream_ptr())
            else:
                rc = lib.cf_int2_decompress(nv.ptr(packed), nv.ptr(u), nv.ptr(v), nv.ptr(base), nv.ptr(recon), n, c,
                                            nv.stream_ptr())
            nv.check(rc, "decompress")
        run.keep = (base, packed, u, v, recon)
        return run

    pay = code + scales
    return [
        Point(f"{codec_name}.compress_ef", n, c, 6 * e + code + scales, 6 * e + code, mk_compress(True), pay),
        Point(f"{codec_name}.compress", n, c, 4 * e + code + scales, 4 * e + code, mk_compress(False), pay),
        Point(f"{codec_name}.decompress", n, c, 4 * e + code + scales, 4 * e + code, make_dec, pay),
    ]


def minmax_points(codec_name, n, c, dev):
    codec = nv.CODEC_INT4 if codec_name == "int4" else nv.CODEC_INT8
    e = n * c
    code = e // 2 if codec_name == "int4" else e
    lib = nv.lib()
    cfn = lib.cf_int4_compress if codec_name == "int4" else lib.cf_int8_compress
    dfn = lib.cf_int4_decompress if codec_name == "int4" else lib.cf_int8_decompress
    ws_bytes = nv.workspace_bytes(codec, n, c)

    def make_c(i):
        x, base = _rand_pair(n, c, dev, i)
        codes = torch.empty(code, dtype=torch.uint8, device=dev)
        s = torch.empty(c, dtype=torch.half, device=dev)
        m = torch.empty(c, dtype=torch.half, device=dev)
        nb = torch.empty_like(x)
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)

        def run():
            nv.check(cfn(nv.ptr(x), nv.ptr(base), nv.ptr(nb), nv.ptr(codes), nv.ptr(s), nv.ptr(m), n, c, nv.ptr(ws),
                         ws.numel(), nv.stream_ptr()), "compress")
        run.keep = (x, base, codes, s, m, nb, ws)
        return run

    def make_d(i):
        x, base = _rand_pair(n, c, dev, i)
        codes = torch.randint(0, 256, (code,), dtype=torch.uint8, device=dev)
        s = (torch.rand(c, device=dev) * 0.1 + 0.01).half()
        m = (-torch.rand(c, device=dev)).half() if codec_name == "int4" else torch.zeros(c, dtype=torch.int16, device=dev)
        recon = torch.empty_like(x)

        def run():
            nv.check(dfn(nv.ptr(codes), nv.ptr(s), nv.ptr(m), nv.ptr(base), nv.ptr(recon), n, c, nv.stream_ptr()), "dec")
        run.keep = (base, codes, s, m, recon)
        return run

    return [
        Point(f"{codec_name}.compress_ef", n, c, 6 * e + code + 4 * c, 6 * e + code, make_c, code + 4 * c),
        Point(f"{codec_name}.decompress", n, c, 4 * e + code + 4 * c, 4 * e + code, make_d, code + 4 * c),
    ]


def topk_points(m, n, c, dev):
    e = n * c
    e -= e % 1024
    lib = nv.lib()
    code = 2 * e // m + e // (2 * m)

    def make_c(i):
        g = torch.Generator(device=dev).manual_seed(i)
        x = torch.randn(e, generator=g, device=dev).half()
        base = (x.float() + 0.3 * torch.randn(e, generator=g, device=dev)).half()
        val = torch.empty(e // m, dtype=torch.half, device=dev)
        idx = torch.empty(e // (2 * m), dtype=torch.uint8, device=dev)
        nb = torch.empty_like(x)

        def run():
            nv.check(lib.cf_topk_compress(nv.ptr(x), nv.ptr(base), nv.ptr(nb), nv.ptr(val), nv.ptr(idx), e, m,
                                          nv.stream_ptr()), "topk")
        run.keep = (x, base, val, idx, nb)
        return run

    def make_d(i):
        g = torch.Generator(device=dev).manual_seed(i)
        base = torch.randn(e, generator=g, device=dev).half()
        val = torch.randn(e // m, generator=g, device=dev).half()
        idx = torch.randint(0, 256, (e // (2 * m),), dtype=torch.uint8, device=dev)
        if m < 16:  # nibbles must be < m
            idx = ((idx >> 4) % m << 4 | (idx & 15) % m).to(torch.uint8)
        recon = torch.empty_like(base)

        def run():
            nv.check(lib.cf_topk_decompress(nv.ptr(val), nv.ptr(idx), nv.ptr(base), nv.ptr(recon), e, m,
                                            nv.stream_ptr()), "topk_dec")
        run.keep = (base, val, idx, recon)
        return run

    return [
        Point(f"topk{m}.compress_ef", n, c, 6 * e + code, 6 * e + code, make_c, code),
        Point(f"topk{m}.decompress", n, c, 4 * e + code, 4 * e + code, make_d, code),
    ]


def lowrank_points(r, n, c, dev, iters=2):
    e = n * c
    lib = nv.lib()
    ws_bytes = nv.workspace_bytes(nv.CODEC_LOWRANK, n, c, r)
    code = 2 * r * (n + c)

    def make_p(i):
        x, base = _rand_pair(n, c, dev, i)
        q0 = torch.linalg.qr(torch.randn(c, r, device=dev))[0].contiguous()
        u = torch.empty(n, r, dtype=torch.half, device=dev)
        v = torch.empty(r, c, dtype=torch.half, device=dev)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)

        def run():
            nv.check(lib.cf_lowrank_project(nv.ptr(x), nv.ptr(base), nv.ptr(q0), nv.ptr(u), nv.ptr(v), None, n, c, r,
                                            iters, nv.ptr(ws), ws.numel(), nv.stream_ptr()), "project")
        run.keep = (x, base, q0, u, v, ws)
        return run

    def make_r(i):
        g = torch.Generator(device=dev).manual_seed(i)
        base = torch.randn(n, c, generator=g, device=dev).half()
        u = torch.randn(n, r, generator=g, device=dev).half()
        v = torch.randn(r, c, generator=g, device=dev).half()
        recon = torch.empty_like(base)

        def run():
            nv.check(lib.cf_lowrank_reconstruct(nv.ptr(u), nv.ptr(v), nv.ptr(base), nv.ptr(recon), n, c, r,
                                                nv.stream_ptr()), "reconstruct")
        run.keep = (base, u, v, recon)
        return run

    # projector: delta is recomputed from x and base in each of the 2*iters+2 passes
    passes = 2 * iters + 2
    return [
        Point(f"lowrank{r}.project", n, c, passes * 4 * e + code, 4 * e, make_p, code),
        Point(f"lowrank{r}.reconstruct", n, c, 4 * e + code, 4 * e, make_r, code),
    ]


def time_point(pt: Point, reps: int, max_sets: int, mem_cap: int):
    sets = max(2, min(max_sets, -(-2 * L2_BYTES // max(pt.footprint, 1)) + 1))
    while sets > 2 and sets * pt.footprint > mem_cap:
        sets -= 1
    runs = [pt.make(i) for i in range(sets)]
    for r in runs:  # warm-up (also sets func attributes outside capture)
        r()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for r in runs:
            r()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / (reps * sets)
    del g, runs
    return us, sets


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes-mb", default="1,8,27,64,256,1024")
    ap.add_argument("--ops", default="binary,int2,int4,int8,topk,lowrank")
    ap.add_argument("--topk-m", default="2,4,8,16")
    ap.add_argument("--ranks", default="4,8,16,32,64")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--max-sets", type=int, default=24)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--md", default="")
    ap.add_argument("--shapes", default="", help="extra NxC shapes, e.g. 4608x3072,576x3072,8192x1152")
    args = ap.parse_args()
    assert torch.cuda.is_available(), "sweep.py needs a CUDA device"
    dev = torch.device("cuda", 0)
    peak, peak_src = hbm_peak()
    shapes = []
    for mb in [float(s) for s in args.sizes_mb.split(",") if s]:
        n = int(mb * (1 << 20) / (2 * C_FIXED))
        n -= n % 8
        shapes.append((max(n, 8), C_FIXED))
    for s in [s for s in args.shapes.split(",") if s]:
        n, c = s.lower().split("x")
        shapes.append((int(n), int(c)))
    ops = args.ops.split(",")
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    rows = []
    with open(args.out, "w") as fout:
        for n, c in shapes:
            pts = []
            for op in ops:
                if op in ("binary", "int2"):
                    pts += sign_points(op, n, c, dev)
                elif op in ("int4", "int8"):
                    pts += minmax_points(op, n, c, dev)
                elif op == "topk":
                    for m in [int(s) for s in args.topk_m.split(",")]:
                        pts += topk_points(m, n, c, dev)
                elif op == "lowrank":
                    for r in [int(s) for s in args.ranks.split(",")]:
                        pts += lowrank_points(r, n, c, dev)
            for pt in pts:
                try:
                    us, sets = time_point(pt, args.reps, args.max_sets, 40 << 30)
                except Exception as ex:  # keep sweeping
                    print(f"# {pt.name} {n}x{c}: {type(ex).__name__}: {ex}", file=sys.stderr)
                    continue
                gbs = pt.algo_bytes / us / 1e3
                row = {"op": pt.name, "N": n, "C": c, "tensor_mb": n * c * 2 / (1 << 20), "us": us, "algo_bytes": pt.algo_bytes,
                       "gbs": gbs, "frac_hbm": gbs / peak, "peak": peak, "peak_src": peak_src, "sets": sets,
                       "payload_bytes": pt.payload_bytes,
                       "payload_us_at_770gbs": (pt.payload_bytes / 770e3) if pt.payload_bytes else None}
                rows.append(row)
                fout.write(json.dumps(row) + "\n")
                fout.flush()
                print(f"{pt.name:24s} {n:7d}x{c:<5d} {us:10.2f} us  {gbs:8.1f} GB/s  {gbs / peak * 100:5.1f}% of {peak_src} HBM")
                torch.cuda.empty_cache()
    if args.md:
        with open(args.md, "w") as f:
            f.write(f"| op | N x C | tensor MB | us/call | algorithmic GB/s | frac of {peak_src} HBM peak ({peak:.0f} GB/s) | payload B | payload us @770 GB/s |\n")
            f.write("|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                f.write(f"| {r['op']} | {r['N']}x{r['C']} | {r['tensor_mb']:.1f} | {r['us']:.2f} | {r['gbs']:.0f} | {r['frac_hbm']:.3f} | "
                        f"{r['payload_bytes']} | {r['payload_us_at_770gbs']:.2f} |\n")


if __name__ == "__main__":
    main()
