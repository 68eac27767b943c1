#!/usr/bin/env python
"""A few steps of the W-rank exchange with W VIRTUAL ranks on one GPU (engine.LocalWorld): the fused put's
kernels store into W receive regions in local memory, which is what a peer mapping looks like to them.  Used
under `ncu --set full` (one process, one GPU -- ncu cannot replay a multi-rank job) to get the DRAM traffic
and stall picture of the PUT kernels at the shard sizes of N = 2 / 4 / 8:

  ncu --set full --clock-control none -k regex:'k_delta_stats|k_finalize|k_apply|k_publish' -s 40 -c 12 \
      -f -o gpurun_out/vr8 python tools/virtual_ranks_step.py --world 8 --layers 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--rows", type=int, default=4608)
    ap.add_argument("--ch", type=int, default=3072)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--codec", default="binary")
    a = ap.parse_args()
    import compactfusion_b200 as cf
    from compactfusion_b200.engine import LocalWorld
    T = cf.COMPACT_COMPRESS_TYPE
    dev = torch.device("cuda:0")
    n = a.rows // a.world
    lw = LocalWorld(a.world, a.layers, n, a.ch, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    x0 = [[[torch.randn(n, a.ch, generator=g, device=dev) for _ in range(a.world)] for _ in range(2)] for _ in range(a.layers)]
    for t in range(a.steps + 1):
        ct = T(a.codec) if t >= 1 else T.WARMUP
        for l in range(a.layers):
            ks = [(x0[l][0][r] + 0.2 * t * torch.randn(n, a.ch, generator=g, device=dev)).half() for r in range(a.world)]
            vs = [(x0[l][1][r] + 0.2 * t * torch.randn(n, a.ch, generator=g, device=dev)).half() for r in range(a.world)]
            lw.exchange_all(l, ks, vs, ct)
    torch.cuda.synchronize()
    assert not any(e.p2p_error() for e in lw.engines)
    print("ok", a.world, n, a.ch)


if __name__ == "__main__":
    main()
