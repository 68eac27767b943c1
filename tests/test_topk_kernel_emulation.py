"""Execute the SOURCE of the SPARSE 1:m ("top-k") kernels (csrc/cf_topk.cu) on the CPU (tests/cuda_emulation.py)
against what the reference's sim_topk produced on the committed input (tests/golden/codecs.npz) and against the
oracle's values / nibble indices: bit-exact, for m = 2, 4, 8, 16, with and without the fused residual / EF."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import cuda_emulation as emu
from conftest import GOLDEN, bits16, h16
from oracle import codecs as oc

RUNNER = r'''
template <int M>
static int run(const __half* x, const __half* base, int64_t numel) {
  using namespace cf;
  const int64_t nt = numel / 32;
  std::vector<__half> val(numel / M), new_base(numel), recon(numel);
  std::vector<uint8_t> idx(numel / (2 * M));
  launch(3, 1, 256, 1, [&] { k_topk_compress<M>(x, base, new_base.data(), val.data(), idx.data(), nt); });
  launch(2, 1, 256, 1, [&] { k_topk_decompress<M>(val.data(), idx.data(), base, recon.data(), nt); });
  fwrite(val.data(), 2, val.size(), stdout);
  fwrite(idx.data(), 1, idx.size(), stdout);
  fwrite(new_base.data(), 2, numel, stdout);
  fwrite(recon.data(), 2, numel, stdout);
  return 0;
}
int main(int argc, char** argv) {  // m x.bin [base.bin]
  const int m = atoi(argv[1]);
  auto x = slurp(argv[2]);
  std::vector<unsigned char> b; if (argc > 3) b = slurp(argv[3]);
  const __half* xh = reinterpret_cast<const __half*>(x.data());
  const __half* bh = b.empty() ? nullptr : reinterpret_cast<const __half*>(b.data());
  const int64_t numel = (int64_t)x.size() / 2;
  switch (m) {
    case 2: return run<2>(xh, bh, numel);
    case 4: return run<4>(xh, bh, numel);
    case 8: return run<8>(xh, bh, numel);
    default: return run<16>(xh, bh, numel);
  }
}
'''


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    src = open(os.path.join(emu.CSRC, "cf_topk.cu")).read()
    kern = re.search(r"(namespace cf \{.*?)static int topk_check", src, flags=re.S).group(1) + "}  // namespace cf\n"
    assert "k_topk_decompress" in kern and "<<<" not in kern and "asm" not in kern
    d = tmp_path_factory.mktemp("topk_emu")
    return emu.build(d, "#define __align__(n) alignas(n)\n" + emu.SHIM_HEAD + emu.common_source() + kern + emu.SLURP + RUNNER), d


def _run(emulator, m, x, base=None):
    exe, d = emulator
    numel = x.numel()
    (d / "x.bin").write_bytes(x.contiguous().numpy().tobytes())
    cmd = [exe, str(m), str(d / "x.bin")]
    if base is not None:
        (d / "b.bin").write_bytes(base.contiguous().numpy().tobytes())
        cmd.append(str(d / "b.bin"))
    r = subprocess.run(cmd, capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    sizes = [2 * numel // m, numel // (2 * m), 2 * numel, 2 * numel]
    assert len(r.stdout) == sum(sizes)
    o, parts = 0, []
    for s in sizes:
        parts.append(r.stdout[o:o + s])
        o += s
    return (np.frombuffer(parts[0], dtype=np.uint16), np.frombuffer(parts[1], dtype=np.uint8),
            np.frombuffer(parts[2], dtype=np.uint16), np.frombuffer(parts[3], dtype=np.uint16))


@pytest.mark.parametrize("m", [2, 4, 8, 16])
def test_topk_kernel_source_matches_reference_golden_and_oracle(emulator, m):
    g = np.load(os.path.join(GOLDEN, "codecs.npz"))
    x = h16(g["topk/x"])                                  # (A, 1024) rows
    val, idx, new_base, recon = _run(emulator, m, x)
    ref = bits16(h16(g[f"topk/sim_m{m}"])).reshape(-1)
    assert np.array_equal(recon, ref), "decompressed values differ from the reference's sim_topk"
    assert np.array_equal(new_base, ref), "fused reconstruction (base == NULL) differs"
    o_val, o_idx = oc.topk_compress(x, m)
    assert np.array_equal(val, bits16(o_val).reshape(-1)) and np.array_equal(idx, o_idx.reshape(-1))
    # fused residual + error feedback: new_base = base + sparsify(x - base) == what the receiver reconstructs
    gen = torch.Generator().manual_seed(m)
    base = (x.float() + 0.3 * torch.randn(x.shape, generator=gen)).half()
    val2, idx2, nb2, rec2 = _run(emulator, m, x, base)
    d = x - base
    o_val2, o_idx2 = oc.topk_compress(d, m)
    assert np.array_equal(val2, bits16(o_val2).reshape(-1)) and np.array_equal(idx2, o_idx2.reshape(-1))
    want = base + oc.topk_decompress(o_val2, o_idx2, m).view(base.shape)
    assert np.array_equal(nb2, bits16(want).reshape(-1)) and np.array_equal(rec2, nb2)


def test_topk_ties_pick_the_lowest_index_in_the_kernel_source(emulator):
    x = torch.zeros(1, 1024, dtype=torch.half)
    x[0, 0:4] = torch.tensor([1.0, -1.0, 1.0, 0.5])      # tie between index 0 and 1 and 2 -> 0
    x[0, 4:8] = torch.tensor([0.25, -2.0, 2.0, 2.0])     # tie between 1, 2, 3 -> 1
    val, idx, _, recon = _run(emulator, 4, x)
    assert idx[0] == ((0 << 4) | 1)
    assert val[0] == 0x3C00 and val[1] == 0xC000          # 1.0, -2.0
