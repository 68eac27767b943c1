#!/usr/bin/env python
"""Per-kernel GPU times of one library call via CUPTI (torch.profiler): seconds instead of an ncu run,
warm caches, real overlap -- use it to find WHERE a call spends its time before profiling one kernel with
`ncu --set full`.  Not a bench value (profiler overhead on launch gaps); kernel durations are accurate.

  python tools/kernel_times.py lowrank --shape 4608x3072 --rank 32 --iters 2
  python tools/kernel_times.py codec --codec int4 --shape 4608x3072
  python tools/kernel_times.py step --codec binary --layers 8 [--overlap]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def kernel_table(fn, reps):
    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        dev_us = getattr(e, "device_time_total", None)
        if dev_us is None:
            dev_us = getattr(e, "cuda_time_total", 0.0)
        if dev_us and e.device_type.name != "CPU":
            rows.append((e.key, e.count, dev_us / max(e.count, 1), dev_us / reps))
    if not rows:  # older / newer torch: kernels are listed among all events
        for e in prof.key_averages():
            dev_us = getattr(e, "self_device_time_total", None) or getattr(e, "self_cuda_time_total", 0.0)
            if dev_us:
                rows.append((e.key, e.count, dev_us / max(e.count, 1), dev_us / reps))
    rows.sort(key=lambda r: -r[3])
    total = sum(r[3] for r in rows)
    print(f"| kernel | launches / call | avg us | us / call | share |\n|---|---:|---:|---:|---:|")
    for name, count, avg, per_call in rows:
        print(f"| `{name[:90]}` | {count / reps:.1f} | {avg:.2f} | {per_call:.2f} | {per_call / total:.1%} |")
    print(f"\nsum of kernel time per call: {total:.1f} us ({reps} calls profiled)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["lowrank", "codec", "step"])
    ap.add_argument("--shape", default="4608x3072")
    ap.add_argument("--rank", type=int, default=32)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--codec", default="binary")
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--overlap", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    n, c = (int(v) for v in a.shape.split("x"))
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, c, generator=g, device=dev).half()
    base = (x.float() + 0.3 * torch.randn(n, c, generator=g, device=dev)).half()
    if a.what == "lowrank":
        from compactfusion_b200.compress_lowrank import lowrank_project, lowrank_reconstruct

        def fn():
            u, v, _ = lowrank_project(x, base, a.rank, a.iters)
            lowrank_reconstruct(u, v, base)
    elif a.what == "codec":
        import compactfusion_b200 as cf
        T = cf.COMPACT_COMPRESS_TYPE
        fast = a.codec in ("binary", "int2")
        cf.compact_init(cf.CompactConfig(enabled=True, residual=1, ef=True, fastpath=fast, comp_rank=-1 if fast else a.rank,
                                         sparse_ratio=8, compress_func=lambda l, s: T(a.codec)))
        cf.compact_compress("0-0-k", x, T.WARMUP, update_cache=True)
        cf.compact_decompress("1-0-k", x, T.WARMUP, tuple(x.shape), update_cache=True)

        def fn():
            p = cf.compact_compress("0-0-k", base, T(a.codec), update_cache=True)
            cf.compact_decompress("1-0-k", p, T(a.codec), tuple(x.shape), update_cache=True)
    else:
        import compactfusion_b200 as cf
        from compactfusion_b200.engine import PatchGatherEngine
        T = cf.COMPACT_COMPRESS_TYPE
        eng = PatchGatherEngine(a.layers, n, c, device=dev)
        ks = [x.clone() for _ in range(a.layers)]
        vs = [base.clone() for _ in range(a.layers)]
        eng.step(ks, vs, T.WARMUP)
        graph = eng.capture_step(vs, ks, T(a.codec), overlap=a.overlap)

        def fn():
            graph.replay()
    kernel_table(fn, a.reps)


if __name__ == "__main__":
    main()
