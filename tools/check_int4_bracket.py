#!/usr/bin/env python
"""Exhaustive-in-the-numerator check (CPU, numpy) of the exact-quotient shortcut of the INT4 encode
(csrc/cf_minmax_codecs.cu, int4_codes2): the reference computes fp16(a / s) = RN16(RN32(a / s)); the kernel
computes t = RN32(a * RN32(1 / s)) and accepts RN16 of the bracket [t (1 - 2^-21), t (1 + 2^-21)] when both ends
round to the same fp16 number (RN16 is monotone and RN32(a / s) lies inside the bracket), else divides.
    python tools/check_int4_bracket.py [n_scales]
Every non-negative finite fp16 numerator x n_scales random fp16 scales (plus the smallest / largest ones)."""
import sys

import numpy as np


def main():
    n_scales = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    rng = np.random.default_rng(0)
    a16 = np.arange(0, 0x7C00, dtype=np.uint16).view(np.float16)            # every finite fp16 >= 0
    s_bits = rng.integers(1, 0x7C00, size=n_scales, dtype=np.uint16)
    s_bits[:8] = [1, 2, 0x03FF, 0x0400, 0x3C00, 0x7BFF, 0x0401, 0x2E66]      # subnormals, 1.0, max, ...
    a = a16.astype(np.float32)
    k_lo, k_hi = np.float32(1.0) - np.float32(2.0 ** -21), np.float32(1.0) + np.float32(2.0 ** -21)
    wrong = fallback = total = 0
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        for sb in s_bits:
            s = np.array([sb], dtype=np.uint16).view(np.float16).astype(np.float32)[0]
            ref = (a / s).astype(np.float16)                                 # RN16(RN32(a / s))
            rcp = np.float32(1.0) / s
            t = a * rcp
            lo, hi = (t * k_lo).astype(np.float16), (t * k_hi).astype(np.float16)
            same = lo.view(np.uint16) == hi.view(np.uint16)
            wrong += int((lo.view(np.uint16)[same] != ref.view(np.uint16)[same]).sum())
            fallback += int((~same).sum())
            total += a.size
    print(f"{total} quotients: {wrong} accepted-but-wrong, {fallback} ({100.0 * fallback / total:.3f} %) sent to the division")
    return 1 if wrong else 0


if __name__ == "__main__":
    sys.exit(main())
