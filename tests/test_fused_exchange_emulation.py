"""The fused compress + one-sided exchange on the CPU, W ranks in one process (kernel SOURCE, tests/
cuda_emulation.py): every rank's codec kernels store K's and V's payload straight into slot (origin) of all W
receive regions and bump the flags (cf_sign_compress_put: k_delta_stats_tma<.., PUT>, k_finalize_scales<.., PUT>,
k_int2_encode_tma<.., PUT>), then every rank reconstructs all W origins in one flag-waiting launch
(k_apply_codes_tma with wait_flag / expected, 2W tensors).  W = 8 never ran on a GPU this round; here the
fan-out tables, slot offsets, flag arithmetic (flag == the publishing kernel's grid size == the sender's count)
and the batched tile schedule at 16 tensors are executed and the result is held bit-identical to the plain
(non-fused) pipelined kernels on the same inputs.  Ranks run one after another, so the device-side wait never
spins: what is checked is the data path and the counters, not the inter-GPU memory ordering.
"""
import os
import re
import subprocess

import pytest
import torch

import cuda_emulation as emu

RUNNER = r'''
using namespace cf;
static PipeArgs args_of(const PipeGeom& g, int rows_per_cta) {
  PipeArgs a{};
  a.TX = g.TX; a.TY = g.TY; a.R = g.R; a.stages = g.stages; a.chunk_rows = g.chunk_rows; a.u_cap = g.u_cap;
  a.tile_bytes = g.tile_bytes; a.stage_bytes = g.stage_bytes; a.rows_per_cta = rows_per_cta; a.l2_hints = 0; a.early_load = 1;
  return a;
}
static TileSched sched_of(const PipeGeom& g, int N, int batch, int want_ctas, int* n_cta) {
  TileSched ts{};
  ts.tiles_per_tensor = (N + g.R - 1) / g.R;
  ts.total_tiles = ts.tiles_per_tensor * batch;
  int ctas = want_ctas > ts.total_tiles ? ts.total_tiles : want_ctas;
  ts.tiles_per_cta = (ts.total_tiles + ctas - 1) / ctas;
  const int cap = g.u_cap / g.R > 0 ? g.u_cap / g.R : 1;
  if (ts.tiles_per_cta > cap) ts.tiles_per_cta = cap;
  *n_cta = (ts.total_tiles + ts.tiles_per_cta - 1) / ts.tiles_per_cta;
  return ts;
}

template <int MODE>
static int run(const std::vector<unsigned char>& xs, const std::vector<unsigned char>& bs, int W, int N, int C) {
  constexpr int per_code = (MODE == MODE_BINARY) ? 8 : 4;
  const size_t E = (size_t)N * C, code_bytes = E / per_code, pay = code_bytes + 2 * (size_t)N + 2 * (size_t)C;
  const size_t pay_al = (pay + 15) / 16 * 16;
  auto X = [&](int r, int j) { return reinterpret_cast<const __half*>(xs.data()) + ((size_t)(2 * r + j)) * E; };
  auto Bp = [&](int r, int j) { return reinterpret_cast<const __half*>(bs.data()) + ((size_t)(2 * r + j)) * E; };
  // receive regions: region[q] = slots (origin r) x {K, V}; flags[q][r]
  std::vector<std::vector<unsigned char>> region(W, std::vector<unsigned char>((size_t)W * 2 * pay_al + 64, 0xCD));
  auto slot = [&](int q, int r, int j) {
    unsigned char* p = region[q].data();
    p += (16 - (reinterpret_cast<uintptr_t>(p) & 15u)) & 15u;
    return p + ((size_t)r * 2 + j) * pay_al;
  };
  std::vector<std::vector<uint32_t>> flags(W, std::vector<uint32_t>(W, 0));
  std::vector<uint32_t> count(W, 0), ticket(W, 0), err(W, 0);
  const PipeGeom g1 = make_pipe_geom(C, 2, 0, false), g2 = make_pipe_geom(C, 2, 0, true), g3 = make_pipe_geom(C, 1, C / per_code, true);
  if (!g1.ok || g1.G != 1 || g1.ctas_per_sm != 2) { fprintf(stderr, "geometry\n"); return 3; }
  uint32_t publish_grid = 0;
  // reference (non-fused) payloads per (rank, tensor): [codes | U | V]
  std::vector<std::vector<unsigned char>> ref(2 * W, std::vector<unsigned char>(pay, 0));
  for (int r = 0; r < W; ++r) {
    int64_t rpc = (N + 1) / 2; rpc = (rpc + g1.R - 1) / g1.R * g1.R;     // 2 row blocks
    const int B = (int)((N + rpc - 1) / rpc);
    std::vector<__half> rowmean(2 * (size_t)N);
    std::vector<float> tokpart(2 * B), colpart(2 * (size_t)B * C);
    for (int fused = 0; fused < 2; ++fused) {
      StatsParams sp{}; FinalizeParams fp{}; FanOut f{};
      sp.N = fp.N = N; sp.C = fp.C = C; sp.rows_per_cta = (int)rpc; fp.B = B;
      f.n_dst = W; f.u_off = code_bytes; f.v_off = code_bytes + 2 * (size_t)N; f.count = &count[r]; f.done = &ticket[r];
      for (int q = 0; q < W; ++q) f.flag[q] = &flags[q][r];
      for (int t = 0; t < 2; ++t) {
        sp.x[t] = X(r, t); sp.base[t] = Bp(r, t);
        sp.rowmean[t] = rowmean.data() + (size_t)t * N; fp.rowmean[t] = sp.rowmean[t];
        sp.tokpart[t] = tokpart.data() + t * B; fp.tokpart[t] = sp.tokpart[t];
        sp.colpart[t] = colpart.data() + (size_t)t * B * C; fp.colpart[t] = sp.colpart[t];
        sp.packed[t] = ref[2 * r + t].data();
        fp.scale_u[t] = reinterpret_cast<__half*>(ref[2 * r + t].data() + f.u_off);
        fp.scale_v[t] = reinterpret_cast<__half*>(ref[2 * r + t].data() + f.v_off);
        for (int q = 0; q < W; ++q) f.dst[t * W + q] = slot(q, r, t);
      }
      const PipeArgs a1 = args_of(g1, (int)rpc);
      emu_ncompute = g1.TX * g1.TY;
      if (fused && MODE == MODE_BINARY) launch(B, 2, g1.TX * g1.TY + 32, 1, [&] { k_delta_stats_tma<MODE, 1, 2, (MODE == MODE_BINARY)>(sp, a1, f); });
      else launch(B, 2, g1.TX * g1.TY + 32, 1, [&] { k_delta_stats_tma<MODE, 1, 2, false>(sp, a1, f); });
      const unsigned fgrid = (C + 31) / 32;
      if (fused) launch(fgrid, 2, 1024, 1, [&] { k_finalize_scales<MODE, true>(fp, f, MODE == MODE_BINARY ? 1 : 0); });
      else launch(fgrid, 2, 1024, 1, [&] { k_finalize_scales<MODE, false>(fp, FanOut{}, 0); });
      if (MODE == MODE_BINARY) publish_grid = fgrid * 2;
      if (MODE == MODE_INT2) {
        Int2EncodeParams ep{}; ep.N = N; ep.C = C;
        for (int t = 0; t < 2; ++t) {
          unsigned char* own = fused ? slot(r, r, t) : ref[2 * r + t].data();
          ep.x[t] = X(r, t); ep.base[t] = Bp(r, t);
          ep.scale_u[t] = reinterpret_cast<const __half*>(own + f.u_off);
          ep.scale_v[t] = reinterpret_cast<const __half*>(own + f.v_off);
          ep.packed[t] = own; ep.new_base[t] = nullptr;
        }
        int n_cta = 0; const TileSched ts = sched_of(g2, N, 2, 3, &n_cta);
        const PipeArgs a2 = args_of(g2, 0);
        emu_ncompute = g2.TX * g2.TY;
        if (fused) { launch(n_cta, 1, g2.TX * g2.TY + 32, 1, [&] { k_int2_encode_tma<1, 2, true>(ep, a2, ts, f); }); publish_grid = n_cta; }
        else launch(n_cta, 1, g2.TX * g2.TY + 32, 1, [&] { k_int2_encode_tma<1, 2, false>(ep, a2, ts, FanOut{}); });
      }
    }
  }
  // counters: every flag and every sender's count advanced by the publishing kernel's grid size; tickets reset
  for (int r = 0; r < W; ++r) {
    if (count[r] != publish_grid || ticket[r] != 0) { fprintf(stderr, "rank %d: count %u (grid %u) ticket %u\n", r, count[r], publish_grid, ticket[r]); return 4; }
    for (int q = 0; q < W; ++q) if (flags[q][r] != publish_grid) { fprintf(stderr, "flag[%d][%d] = %u\n", q, r, flags[q][r]); return 4; }
  }
  // every slot on every rank == the non-fused payload of that origin, byte for byte
  for (int q = 0; q < W; ++q) for (int r = 0; r < W; ++r) for (int t = 0; t < 2; ++t)
    if (memcmp(slot(q, r, t), ref[2 * r + t].data(), pay) != 0) { fprintf(stderr, "slot (dst %d, origin %d, tensor %d) differs from the plain payload\n", q, r, t); return 5; }
  // flag-waiting reconstruct of all W origins x {K, V} on rank q = 1, against per-tensor plain applies
  const int q = W > 1 ? 1 : 0;
  std::vector<std::vector<__half>> out(2 * W, std::vector<__half>(E)), want(2 * W, std::vector<__half>(E));
  ApplyParams ap{}; ap.N = N; ap.C = C; ap.K = 1; ap.expected = &count[q]; ap.error = &err[q]; ap.wait_mode = 0;
  for (int r = 0; r < W; ++r) for (int t = 0; t < 2; ++t) {
    const int i = 2 * r + t;
    ap.packed[i] = slot(q, r, t);
    ap.scale_u[i] = reinterpret_cast<const __half*>(slot(q, r, t) + code_bytes);
    ap.scale_v[i] = reinterpret_cast<const __half*>(slot(q, r, t) + code_bytes + 2 * (size_t)N);
    ap.base[i] = Bp(r, t); ap.recon[i] = out[i].data(); ap.wait_flag[i] = &flags[q][r];
  }
  { int n_cta = 0; const TileSched ts = sched_of(g3, N, 2 * W, 5, &n_cta); const PipeArgs a3 = args_of(g3, 0);
    emu_ncompute = g3.TX * g3.TY;
    launch(n_cta, 1, g3.TX * g3.TY + 32, 1, [&] { k_apply_codes_tma<MODE, 1, 2>(ap, a3, ts); }); }
  for (int i = 0; i < 2 * W; ++i) {
    ApplyParams a1{}; a1.N = N; a1.C = C; a1.K = 1;
    a1.packed[0] = ref[i].data(); a1.scale_u[0] = reinterpret_cast<const __half*>(ref[i].data() + code_bytes);
    a1.scale_v[0] = reinterpret_cast<const __half*>(ref[i].data() + code_bytes + 2 * (size_t)N);
    a1.base[0] = Bp(i / 2, i % 2); a1.recon[0] = want[i].data();
    int n_cta = 0; const TileSched ts = sched_of(g3, N, 1, 2, &n_cta); const PipeArgs a3 = args_of(g3, 0);
    emu_ncompute = g3.TX * g3.TY;
    launch(n_cta, 1, g3.TX * g3.TY + 32, 1, [&] { k_apply_codes_tma<MODE, 1, 2>(a1, a3, ts); });
    if (memcmp(out[i].data(), want[i].data(), E * 2) != 0) { fprintf(stderr, "reconstruction of tensor %d differs\n", i); return 6; }
  }
  if (err[q] != 0) { fprintf(stderr, "wait error word set\n"); return 7; }
  printf("FUSED_EXCHANGE_OK W=%d grid=%u\n", W, publish_grid);
  return 0;
}

int main(int argc, char** argv) {  // <binary|int2> x.bin base.bin W N C   (x.bin: W ranks x {K, V} x (N, C) fp16)
  const std::string mode = argv[1];
  auto x = slurp(argv[2]), b = slurp(argv[3]);
  const int W = atoi(argv[4]), N = atoi(argv[5]), C = atoi(argv[6]);
  return mode == "binary" ? run<MODE_BINARY>(x, b, W, N, C) : run<MODE_INT2>(x, b, W, N, C);
}
'''


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    src = open(os.path.join(emu.CSRC, "cf_sign_codecs.cu")).read()
    kern = re.search(r"(namespace cf \{.*?\n\}  // namespace cf\n)", src, flags=re.S).group(1)
    # the two PTX statements of the publish path get their host meaning instead of being dropped
    red = 'asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");'
    bar = 'asm volatile("bar.sync 1, %0;" ::"r"(ncompute) : "memory");'
    assert red in kern and bar in kern
    kern = kern.replace(red, "__atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);").replace(bar, "compute_sync(ncompute);")
    kern = emu.strip_asm(kern).replace("extern __shared__ float smem[];", "float* smem = emu_smem;")
    tma = open(os.path.join(emu.CSRC, "cf_sign_tma.cuh")).read()
    tma = tma.replace('#pragma once', '').replace('#include "cf_pipe.cuh"', '')
    for kind in ("relaxed", "acquire"):   # flag polls: plain volatile loads
        ptx = f'asm volatile("ld.{kind}.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");'
        assert ptx in tma
        tma = tma.replace(ptx, "v = *reinterpret_cast<const volatile uint32_t*>(p);")
    tma = emu.strip_asm(tma).replace("extern __shared__ __align__(128) unsigned char pipe_smem_raw[];",
                                     "unsigned char* pipe_smem_raw = emu_smem_bytes;")
    hook = "#define EMU_LAUNCH_HOOK if (emu_ncompute > 0) pthread_barrier_init(&compute_bar, nullptr, emu_ncompute);\n"
    fwd = "static pthread_barrier_t compute_bar; static int emu_ncompute;\n"
    pipe_shim = emu.PIPE_SHIM.replace("static pthread_barrier_t compute_bar;\n", "").replace(
        "static int emu_ncompute = 0;   // set by the runner before a launch of a pipelined kernel\n", "")
    d = tmp_path_factory.mktemp("fused_emu")
    text = ("#include <pthread.h>\n" + fwd + hook + emu.SHIM_HEAD + pipe_shim + emu.common_source() + kern + tma
            + emu.SLURP + RUNNER)
    return emu.build(d, text), d


@pytest.mark.parametrize("mode", ["binary", "int2"])
@pytest.mark.parametrize("world,n,c", [(8, 72, 256), (2, 150, 512)])
def test_fused_put_and_flag_waiting_reconstruct_on_cpu(emulator, mode, world, n, c):
    exe, d = emulator
    g = torch.Generator().manual_seed(world * 100 + n)
    x = torch.randn(world * 2, n, c, generator=g).half()
    base = (0.97 * x.float() + 0.2 * torch.randn(world * 2, n, c, generator=g)).half()
    (d / "x.bin").write_bytes(x.numpy().tobytes())
    (d / "b.bin").write_bytes(base.numpy().tobytes())
    r = subprocess.run([exe, mode, str(d / "x.bin"), str(d / "b.bin"), str(world), str(n), str(c)],
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and f"FUSED_EXCHANGE_OK W={world}" in r.stdout, r.stdout + r.stderr[-2000:]


WAIT_RUNNER = r'''
#include <chrono>
// A late sender: the receiver's flag-waiting reconstruct is launched FIRST; another thread then writes the
// payload into the slot and only afterwards raises the flag.  The kernel must not read the slot early.
int main(int argc, char** argv) {   // x.bin base.bin N C
  using namespace cf;
  auto x = slurp(argv[1]), b = slurp(argv[2]);
  const int N = atoi(argv[3]), C = atoi(argv[4]);
  const size_t E = (size_t)N * C, code_bytes = E / 8;
  const __half* xh = reinterpret_cast<const __half*>(x.data());
  const __half* bh = reinterpret_cast<const __half*>(b.data());
  // the payload the sender will deliver (plain pipelined compress)
  std::vector<unsigned char> pay(code_bytes + 2 * (size_t)N + 2 * (size_t)C);
  const PipeGeom g1 = make_pipe_geom(C, 2, 0, false), g3 = make_pipe_geom(C, 1, C / 8, true);
  auto args = [](const PipeGeom& g, int rpc) { PipeArgs a{}; a.TX = g.TX; a.TY = g.TY; a.R = g.R; a.stages = g.stages;
    a.chunk_rows = g.chunk_rows; a.u_cap = g.u_cap; a.tile_bytes = g.tile_bytes; a.stage_bytes = g.stage_bytes;
    a.rows_per_cta = rpc; a.early_load = 1; return a; };
  int64_t rpc = (N + g1.R - 1) / g1.R * g1.R;
  std::vector<__half> rowmean(N); std::vector<float> tokpart(1), colpart(C);
  StatsParams sp{}; sp.x[0] = xh; sp.base[0] = bh; sp.packed[0] = pay.data(); sp.rowmean[0] = rowmean.data();
  sp.tokpart[0] = tokpart.data(); sp.colpart[0] = colpart.data(); sp.N = N; sp.C = C; sp.rows_per_cta = (int)rpc;
  const PipeArgs a1 = args(g1, (int)rpc);
  emu_ncompute = g1.TX * g1.TY;
  launch(1, 1, g1.TX * g1.TY + 32, 1, [&] { k_delta_stats_tma<MODE_BINARY, 1, 2, false>(sp, a1, FanOut{}); });
  FinalizeParams fp{}; fp.rowmean[0] = rowmean.data(); fp.tokpart[0] = tokpart.data(); fp.colpart[0] = colpart.data();
  fp.scale_u[0] = reinterpret_cast<__half*>(pay.data() + code_bytes);
  fp.scale_v[0] = reinterpret_cast<__half*>(pay.data() + code_bytes + 2 * (size_t)N); fp.N = N; fp.C = C; fp.B = 1;
  launch((C + 31) / 32, 1, 1024, 1, [&] { k_finalize_scales<MODE_BINARY, false>(fp, FanOut{}, 0); });
  // receiver: slot full of garbage, flag behind the expected count
  std::vector<unsigned char> slot_mem(pay.size() + 32, 0x5A);
  unsigned char* slot = slot_mem.data() + ((16 - (reinterpret_cast<uintptr_t>(slot_mem.data()) & 15u)) & 15u);
  uint32_t flag = 6, expected = 7, err = 0;
  std::vector<__half> got(E), want(E);
  ApplyParams ap{}; ap.N = N; ap.C = C; ap.K = 1; ap.packed[0] = slot;
  ap.scale_u[0] = reinterpret_cast<const __half*>(slot + code_bytes);
  ap.scale_v[0] = reinterpret_cast<const __half*>(slot + code_bytes + 2 * (size_t)N);
  ap.base[0] = bh; ap.recon[0] = got.data(); ap.wait_flag[0] = &flag; ap.expected = &expected; ap.error = &err;
  const PipeArgs a3 = args(g3, 0);
  TileSched ts{}; ts.tiles_per_tensor = (N + g3.R - 1) / g3.R; ts.total_tiles = ts.tiles_per_tensor; ts.tiles_per_cta = ts.total_tiles;
  std::thread sender([&] {
    std::this_thread::sleep_for(std::chrono::milliseconds(150));
    memcpy(slot, pay.data(), pay.size());
    __sync_synchronize();
    __atomic_store_n(&flag, 7u, __ATOMIC_SEQ_CST);
  });
  emu_ncompute = g3.TX * g3.TY;
  launch(1, 1, g3.TX * g3.TY + 32, 1, [&] { k_apply_codes_tma<MODE_BINARY, 1, 2>(ap, a3, ts); });
  sender.join();
  ApplyParams rf = ap; rf.packed[0] = pay.data(); rf.scale_u[0] = fp.scale_u[0]; rf.scale_v[0] = fp.scale_v[0];
  rf.recon[0] = want.data(); rf.expected = nullptr; rf.wait_flag[0] = nullptr;
  launch(1, 1, g3.TX * g3.TY + 32, 1, [&] { k_apply_codes_tma<MODE_BINARY, 1, 2>(rf, a3, ts); });
  if (err != 0 || memcmp(got.data(), want.data(), E * 2) != 0) { fprintf(stderr, "waited reconstruct differs (err=%u)\n", err); return 1; }
  printf("WAIT_OK\n");
  return 0;
}
'''


def test_flag_waiting_reconstruct_does_not_read_the_slot_before_the_flag(emulator, tmp_path_factory):
    """A late sender (payload written 150 ms after the receiver's kernel started, flag raised last): warp 0
    polls, the row scales / column fragments / code tiles are only fetched behind the wait, the result equals
    the plain reconstruct.  (Early base-tile prefetch before the wait is allowed: bases are local.)"""
    _, d0 = emulator
    src = (d0 / "emu.cpp").read_text()
    src = src[:src.index("using namespace cf;\nstatic PipeArgs args_of")] + WAIT_RUNNER   # same prelude, other main
    d = tmp_path_factory.mktemp("wait_emu")
    exe = emu.build(d, src, name="wait")
    g = torch.Generator().manual_seed(5)
    n, c = 150, 256
    x = torch.randn(n, c, generator=g).half()
    base = (0.97 * x.float() + 0.2 * torch.randn(n, c, generator=g)).half()
    (d / "x.bin").write_bytes(x.numpy().tobytes())
    (d / "b.bin").write_bytes(base.numpy().tobytes())
    r = subprocess.run([exe, str(d / "x.bin"), str(d / "b.bin"), str(n), str(c)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "WAIT_OK" in r.stdout, r.stdout + r.stderr[-2000:]
