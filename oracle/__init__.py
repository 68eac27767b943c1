"""CPU oracle for the CompactFusion residual-compression hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, on the CPU, the arithmetic of
the reference's `xfuser/compact` codecs and residual / error-feedback state
machine.  It exists so that the CUDA kernels in `compactfusion_b200/csrc` can be
checked for parity, and so that `bench.py` can time a CPU baseline.

Who may import it: `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs -- and there only as the checker or the
timed CPU arm.  The product package `compactfusion_b200` never imports it and
has no CPU fallback: it raises if the CUDA extension is missing.

Parity pinning: the oracle is pinned against outputs of the reference itself,
imported in the build container from /root/reference by `oracle/ref_loader.py`
and run eagerly on CPU (`oracle/make_goldens.py` -> `tests/golden/*.npz`,
checked by `tests/test_oracle_golden.py`).  The reference stores no golden
vectors of its own.  Two pieces are pinned differently:
  * the Triton fastpath kernels (xfuser/compact/fastpath.py) cannot run on CPU;
    the oracle is pinned against their `sim_*` twins' building blocks
    (`sim_binary`, eager `quantize_int2`/`dequantize_int2`), which the
    reference's own tests treat as ground truth
    (tests/compact/compress_fastpath_test.py:49-166);
  * the LSE merge of ring attention lives in the third-party `yunchang`
    package (absent; setup.py pins only `yunchang>=0.6.0`): parity unpinned at
    that boundary, we pin it against single-device attention instead.
"""
