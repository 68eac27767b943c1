#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (BASELINE.json: "compress+exchange GB/s and
per-step latency, FLUX 1024^2 at 1/2/4/8 B200").

Workload ("flux1024_patch_parallel", BASELINE.json configs[1]): one denoising step of
FLUX.1-dev at 1024^2 = 57 attention layers x {K, V}, sequence 4096 image + 512 text tokens,
C = 24 x 128 = 3072, fp16.  Each of the N ranks owns 4608/N tokens.  Per layer and step:
residual-compress the local K and V shard against the cached base (1-bit sign codes +
token x channel scales; `--codec int2` for the 2-bit preset), all-gather the compressed
payloads over NCCL, reconstruct every origin's shard (error-feedback cache update) into the
global K/V buffers attention reads.  Synthetic AR(1) activations with per-channel log-normal
scales (no network: no real weights / prompts); step 0 is the reference's uncompressed
WARMUP step and is not timed.

`--workload` selects the other BASELINE.json configs as extra bench lines (same metric, same
kernels): cogvideox5b_ring (configs[2]: 42 layers, bs 2 x 17552 tokens, compressed ring attention --
origins consumed hop by hop, engine.RingExchangeEngine), pixart_patch_parallel / sd3_patch_parallel
(configs[3]: bs 2 x 4096 tokens, C = 1152 / 1536).

  python bench.py [--gpus N --steps K --warmup W]            (torchrun for N > 1)
  python bench.py --impl reference ...                        CPU arm (oracle port of the
                                                              reference's eager torch path)
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

LAYERS, SEQ, CH = 57, 4096 + 512, 3072
METRIC = "compress+exchange GB/s (raw fp16 K/V bytes reconstructed per second, all ranks), FLUX 1024^2 patch parallel"
UNIT = "GB/s"
WORKLOAD, MODE = "flux1024_patch_parallel", "patch"
# BASELINE.json configs.  rows = bs * tokens of the GLOBAL sequence; every rank owns rows / N.
# The default (configs[1], the one `metric` is quoted on) is the headline; the others are extra bench lines.
WORKLOADS = {
    # FLUX.1-dev 1024^2: 57 layers, 4096 image + 512 text tokens, 24 x 128 channels, patch-parallel all-gather
    "flux1024_patch_parallel": dict(layers=57, rows=4096 + 512, ch=3072, mode="patch", bs=1, heads=24,
                                    what="FLUX 1024^2 patch parallel"),
    # CogVideoX-5b 49 frames 720x480: 42 layers, bs 2 (CFG) x 17550 tokens (padded to 17552), 48 x 64 channels,
    # compressed ring attention (configs[2])
    "cogvideox5b_ring": dict(layers=42, rows=2 * 17552, ch=3072, mode="ring", bs=2, heads=48,
                             what="CogVideoX-5b 49x720x480 compressed ring attention"),
    # PixArt-alpha / SD3-medium 1024^2: bs 2 x 4096 tokens, 16 x 72 / 24 x 64 channels, patch parallel (configs[3])
    "pixart_patch_parallel": dict(layers=28, rows=2 * 4096, ch=1152, mode="patch", bs=2, heads=16,
                                  what="PixArt-alpha 1024^2 patch parallel"),
    "sd3_patch_parallel": dict(layers=24, rows=2 * 4096, ch=1536, mode="patch", bs=2, heads=24,
                               what="SD3-medium 1024^2 patch parallel"),
    # BASELINE configs[0]: residual compress / decompress round trip, INT4 + error feedback, one 4096 x 3072 tensor,
    # 28 steps, world size 1 (the reference's CPU-runnable case; here on the GPU through the plugin API)
    "config1_int4_roundtrip": dict(layers=1, rows=4096, ch=3072, mode="roundtrip", bs=1, heads=24,
                                   what="INT4 residual round trip with error feedback, 4096x3072, 28 steps"),
}


def select_workload(name, layers=None):
    """Point the module-level shape constants at `name` (the default leaves them untouched)."""
    global LAYERS, SEQ, CH, METRIC, WORKLOAD, MODE
    w = WORKLOADS[name]
    WORKLOAD, MODE = name, w["mode"]
    LAYERS, SEQ, CH = w["layers"], w["rows"], w["ch"]
    METRIC = ("compress+exchange GB/s (raw fp16 K/V bytes reconstructed per second, all ranks), " + w["what"])
    return layers if layers is not None else LAYERS


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--codec", default="binary", choices=["binary", "int2", "int4", "raw", "lowrank8", "lowrank32", "lowrankq32"],
                   help="raw = the uncompressed exchange of the same K/V (NCCL all-gather of fp16 shards, what "
                        "xDiT does without the plugin): a comparison line, none of our kernels run")
    p.add_argument("--raw-exchange", default="allgather", choices=["allgather", "ring", "async"],
                   help="--codec raw: which uncompressed exchange to time (sync all-gather, NCCL P2P ring relay, "
                        "DistriFusion stale-async all-gather)")
    p.add_argument("--workload", default="flux1024_patch_parallel", choices=sorted(WORKLOADS))
    p.add_argument("--layers", type=int, default=None, help="default: the workload's layer count")
    p.add_argument("--api", default="engine", choices=["engine", "dropin"],
                   help="engine: the whole-step runtime (engine.PatchGatherEngine / RingExchangeEngine, one CUDA graph "
                        "per step); dropin: the reference's own hook, compact_fwd per attention layer (hybrid/"
                        "attn_layer.py:59-64), eager launches, with the attention that follows the exchange stubbed out")
    p.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    p.add_argument("--transport", default="auto", choices=["auto", "p2p", "nccl"],
                   help="payload exchange for N > 1: one-sided NVLink puts (p2p) or NCCL all-gather")
    p.add_argument("--overlap", action="store_true",
                   help="two-chain step: compress (+put) of layer l+1 runs beside the reconstruct of layer l "
                        "(engine._step_overlapped; opt-in until measured)")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-parity", action="store_true", help="skip the oracle parity leg (one extra step of layer 0 "
                   "checked by the CPU oracle on rank 0, outside the timed region)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-gpu-reference", action="store_true",
                   help="skip timing the unmodified reference's GPU kernels (baseline/_ref) in the same run (N = 1)")
    p.add_argument("--cpu-sample-layers", type=int, default=1)
    p.add_argument("--hang-dump", type=float, default=0.0,
                   help="debug: dump all Python stacks and exit if the run takes longer than this many seconds")
    return p.parse_args()


def job_bytes(layers, world):
    """Raw fp16 K/V bytes reconstructed per step, summed over ranks."""
    return world * layers * 2 * SEQ * CH * 2


def synth_activations(n_local, layers, versions, device, rank):
    """AR(1) K/V per layer: x_{t+1} = rho x_t + sqrt(1-rho^2) sigma eps, rho = 0.97, per-channel
    log-normal sigma (log-std 0.355, mean |x| ~ 0.92: the statistics of the reference's
    activation dump)."""
    rho = 0.97
    out = []
    for layer in range(layers):
        per_kv = []
        for j in range(2):
            g = torch.Generator(device=device).manual_seed(1234 + 1000 * rank + 2 * layer + j)
            sigma = torch.exp(0.355 * torch.randn(CH, generator=g, device=device)) * 1.15
            x = torch.randn(n_local, CH, generator=g, device=device) * sigma
            vers = [x.half()]
            for _ in range(versions - 1):
                x = rho * x + (1 - rho * rho) ** 0.5 * sigma * torch.randn(n_local, CH, generator=g, device=device)
                vers.append(x.half())
            per_kv.append(vers)
        out.append(per_kv)
    return out  # out[layer][kv][version]


class ClockSampler:
    """SM clock and throttle reasons of this rank's GPU sampled DURING the timed region: in-process NVML (the
    library behind nvidia-smi; a 20 ms polling thread, initialised before the region starts so that no process
    start-up and no NVML attach to every GPU of the box lands inside it -- on an 8-GPU box a freshly started
    `nvidia-smi -lms` stalled the eager launches of all ranks), falling back to `nvidia-smi -lms 200`."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []
        self.nvml, self.handle, self.samples, self._stop = None, None, [], threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                phys = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() else gpu_index
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample(self):
        n = self.nvml
        try:
            mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            try:
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            self.samples.append((mhz, mask))
        except Exception:
            pass

    def _poll(self):
        while not self._stop.is_set():
            self._sample()
            self._stop.wait(0.02)

    def start(self):
        if self.nvml is not None:
            self._sample()
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._sample()   # the timed region has just ended: still under load
            self._stop.set()
            self.thread.join(timeout=1)
            n, reasons = self.nvml, set()
            bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            for _, mask in self.samples:
                reasons.update(name for name, bit in bits.items() if mask & bit)
            sm = [m for m, _ in self.samples]
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(reasons), "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, codec, n_local, world):
    """DRAM bytes (read + write) per launch of `kernel` from the committed `ncu --set full`
    summary (profiles/traffic.json), or None if that shape was not captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(f"{kernel}|{codec}|n{n_local}|w{world}")
    except Exception:
        return None


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's own eager-torch path (baseline/_ref), else the oracle port
# --------------------------------------------------------------------------------------------
REF_ROOT = os.path.join(ROOT, "baseline", "_ref")


def _load_reference_main():
    """The UNMODIFIED reference's `xfuser.compact.main` from baseline/_ref (staged by tools/stage_reference.sh in
    the build container, git-ignored, travels with gpurun), imported through oracle/ref_loader.py's namespace
    stub; None if it is not staged.  Eager (TORCHDYNAMO_DISABLE=1): the semantics its own tests pin."""
    if not os.path.isdir(os.path.join(REF_ROOT, "xfuser", "compact")):
        return None
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
    os.environ["CF_REFERENCE_ROOT"] = REF_ROOT
    try:
        from oracle import ref_loader
        ref_loader.REFERENCE_ROOT = REF_ROOT
        ref_loader.load_reference()
        import xfuser.compact.main as cm
        from xfuser.compact.utils import COMPACT_COMPRESS_TYPE as RT
        from xfuser.compact.utils import CompactConfig as RCfg
        return cm, RT, RCfg
    except Exception as e:  # noqa: BLE001 -- an unusable staging falls back to the port, and says so
        print(f"[bench] baseline/_ref present but not importable ({type(e).__name__}: {e}); using the oracle port",
              file=sys.stderr)
        return None


class CpuSample:
    """The whole job's work for `sample_layers` layers on the host: for EVERY rank q of `world`, compress q's own
    K and V shard (no cache update) and decompress all `world` origins (cache update) -- compact_all_gather
    without the collective (main.py:390-420).  kind "reference": the reference's own compact_compress /
    compact_decompress with simulate=True (BINARY / INT2 fastpath kernels are Triton-only; simulate routes to
    its eager `sim_binary` / `sim_int2`, main.py:116-127), imported from baseline/_ref.  kind "port": the oracle.
    Inputs and the warm-up step are prepared once; `step()` is the timed unit."""

    def __init__(self, codec, world, sample_layers):
        self.codec, self.world, self.layers = codec, world, sample_layers
        self.n = SEQ // world
        torch.set_num_threads(os.cpu_count() or 1)
        g = torch.Generator().manual_seed(0)
        ref = _load_reference_main()
        self.kind = "reference" if ref else "port"
        lowrank = codec.startswith("lowrank")
        rank_r = int(codec.lstrip("lowrankq")) if lowrank else -1
        value = ("low-rank-int4" if codec.startswith("lowrankq") else "low-rank") if lowrank else codec
        self.xs = []  # [version][tensor][rank]
        x0 = [[torch.randn(self.n, CH, generator=g) for _ in range(world)] for _ in range(sample_layers * 2)]
        for ver in range(2):
            self.xs.append([[(x if ver == 0 else 0.97 * x + 0.243 * torch.randn(self.n, CH, generator=g)).half()
                             for x in per] for per in x0])
        if ref:
            cm, RT, RCfg = ref
            self.cm, self.ct, self.warm = cm, RT(value), RT.WARMUP
            cm.compact_init(RCfg(enabled=True, residual=1, ef=True, simulate=True, comp_rank=rank_r,
                                 compress_func=lambda l, s: RT(value)))
            self._compress = lambda key, x: cm.compact_compress(key, x, self.ct, update_cache=False)
            self._decompress = lambda key, p, shape: cm.compact_decompress(key, p, self.ct, shape, update_cache=True)
            warm = lambda key, x: cm.compact_decompress(key, x, self.warm, x.shape, update_cache=True)  # noqa: E731
        else:
            from oracle.state import OracleCompact
            fast = codec in ("binary", "int2")
            oc = OracleCompact(residual=1, ef=True, fastpath=fast, simulate=(codec == "int4"), comp_rank=rank_r)
            self._compress = lambda key, x: oc.compress(key, x, value, update_cache=False)
            self._decompress = lambda key, p, shape: oc.decompress(key, p, value, shape, update_cache=True)
            warm = lambda key, x: oc.decompress(key, x, "warmup", x.shape, update_cache=True)  # noqa: E731
        # warm-up step: receiver q's base for every origin r
        for i, per in enumerate(self.xs[0]):
            for q in range(world):
                for r in range(world):
                    warm(self._key(i, q, r), per[r].clone())
        self.ver = 1

    def _key(self, tensor, receiver, origin):
        # the reference's key format "{layer}-k-{origin}" (main.py:399,412; utils.py parses int(key.split('-')[0])):
        # the receiving rank is folded into the layer index, since one process holds every rank's cache here
        return f"{tensor * self.world + receiver}-k-{origin}"

    def sample_bytes(self):
        """Raw fp16 K/V bytes reconstructed by one sample step, all ranks (the metric's numerator)."""
        return self.world * self.layers * 2 * SEQ * CH * 2

    def step(self):
        """Seconds for one compressed step of the sampled layers, all `world` ranks' shares, back to back."""
        xs = self.xs[self.ver]
        t0 = time.perf_counter()
        for i, per in enumerate(xs):
            payloads = [self._compress(self._key(i, q, q), per[q]) for q in range(self.world)]
            for q in range(self.world):
                for r in range(self.world):
                    self._decompress(self._key(i, q, r), payloads[r], per[r].shape)
        dt = time.perf_counter() - t0
        self.ver ^= 1
        return dt

    def describe(self, n_steps, total_s):
        src = ("the reference's own compact_compress / compact_decompress (simulate=True, eager torch, unmodified, "
               "baseline/_ref)") if self.kind == "reference" else "oracle port of the reference's eager torch path"
        return (f"{src} on {torch.get_num_threads()} host threads: each step = {self.layers} of {LAYERS} layers "
                f"(K and V) of the {WORKLOAD} step, the shares of all {self.world} rank(s) run back to back "
                f"(per rank: compress own {self.n}x{CH} shard + decompress {self.world} origins); "
                f"{n_steps} steps, {total_s:.1f} s of CPU work; the rate is per byte, not extrapolated")


def cpu_baseline(codec, world, layers, sample_layers, budget_s=12.0):
    """Bounded sample: repeat the sampled layers for ~budget_s of CPU work, keep the mean rate."""
    sample = CpuSample(codec, world, sample_layers)
    sample.step()  # page-in / thread-pool warm-up
    times = []
    while sum(times) < budget_s and len(times) < 500:
        times.append(sample.step())
    dt = sum(times) / len(times)
    return {"value": sample.sample_bytes() / dt / 1e9, "unit": UNIT, "cores": os.cpu_count(), "kind": sample.kind,
            "threads": torch.get_num_threads(), "sample_ms_per_step": dt * 1e3,
            "sample": sample.describe(len(times), sum(times))}


def run_reference(args, world, rank):
    """`--impl reference`: the reference's CPU implementation of the path on the box's host cores (rank 0 only).
    `ms_per_step` is the measured time of one SAMPLE step (so steps x ms_per_step is this run's real duration);
    `value` is the sample's bytes over that time -- a rate, comparable with the GPU arm's."""
    if rank != 0:
        return
    assert args.codec != "raw", "--codec raw is a GPU comparison line; the CPU arm runs the compressed path"
    cpu = CpuSample(args.codec, world, args.cpu_sample_layers)
    for _ in range(args.warmup):
        cpu.step()
    per = [cpu.step() for _ in range(args.steps)]
    dt = sum(per) / len(per)
    val = cpu.sample_bytes() / dt / 1e9
    sample = cpu.describe(args.steps, sum(per))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "exchange": MODE, "codec": args.codec, "layers": args.layers, "seq": SEQ,
                   "channels": CH, "world": world, "shard_rows": SEQ // world, "launch_mode": "cpu (no GPU work)",
                   "schedule": "serial", "transport": "in-process (all ranks' shares on this host)", "l2": "n/a",
                   "sampled_layers": args.cpu_sample_layers,
                   "ms_per_full_step_extrapolated": dt * 1e3 * args.layers / args.cpu_sample_layers},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": cpu.kind,
                         "threads": torch.get_num_threads(), "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
_T0 = time.time()


def note(rank, msg):
    """progress marker on stderr (CF_BENCH_VERBOSE=1): where a multi-rank run is, with wall time"""
    if os.environ.get("CF_BENCH_VERBOSE", "0") == "1":
        print(f"[bench r{rank} +{time.time() - _T0:6.1f}s] {msg}", file=sys.stderr, flush=True)


def p2p_probe(engine_cls, n_local, ctype, device, comp_rank=None):
    """N > 1: run a 2-layer WARMUP + compressed step over the one-sided transport and check that no
    device-side flag wait timed out, BEFORE the 57-layer engine is built on it.  A transport that does
    not deliver would otherwise spin ~2 s in every reconstruct launch of the timed region.  The verdict
    is all-reduced (MIN), so every rank takes the same decision.  Returns (ok, reason)."""
    from compactfusion_b200.utils import COMPACT_COMPRESS_TYPE as T
    ok, why = 1, ""
    try:
        probe = engine_cls(2, n_local, CH, group=None, device=device, transport="auto", comp_rank=comp_rank)
        if probe.prepare(ctype) != "p2p":
            ok, why = 0, "p2p setup failed (CUDA IPC)"
        else:
            g = torch.Generator(device=device).manual_seed(7 + dist.get_rank())
            xs = [[torch.randn(n_local, CH, generator=g, device=device).half() for _ in range(2)] for _ in range(2)]
            probe.step(xs[0], xs[0], T.WARMUP)
            for _ in range(2):  # twice: the second step reuses every slot and flag
                probe.step(xs[1], xs[1], ctype)
            torch.cuda.synchronize()
            if probe.p2p_error():
                ok, why = 0, "a device-side flag wait timed out in the probe step"
            dist.barrier()
            probe.close()  # unmap the probe's regions (cf_ipc_close / cf_ipc_free)
    except Exception as e:  # noqa: BLE001 -- any failure means: do not use this transport
        ok, why = 0, f"{type(e).__name__}: {e}"
    t = torch.tensor([ok], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item()), why


def measure_fidelity(eng, ks_last, vs_last, rank, world, device):
    """Fidelity of what the timed steps produced, over EVERY layer's K and V: this rank's shard of the global
    buffers (the reconstruction all ranks hold of it after the last timed step) against its raw input, reduced
    on the GPU by `cf_error_stats` (one pass per tensor, no host sync until the single read-back) and then over
    ranks (sums added, maxima maxed).  Non-finite figures mean the run is invalid."""
    import math
    from compactfusion_b200.quality import error_stats_raw
    layers = eng.layers
    table = torch.zeros((2 * layers, 4), dtype=torch.float32, device=device)
    for l in range(layers):
        error_stats_raw(eng._shard(eng.global_k[l], rank), ks_last[l].reshape(eng.n, eng.c), out=table[2 * l])
        error_stats_raw(eng._shard(eng.global_v[l], rank), vs_last[l].reshape(eng.n, eng.c), out=table[2 * l + 1])
    t64 = table.double()
    sums = torch.stack([t64[:, 0].sum(), t64[:, 1].sum()])
    maxs = torch.stack([t64[:, 2].max(), t64[:, 3].max()])
    worst_rel = torch.sqrt(t64[:, 0] / t64[:, 1].clamp_min(1e-30)).max().reshape(1)
    finite = torch.isfinite(t64).all().to(torch.float64).reshape(1)
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
        dist.all_reduce(worst_rel, op=dist.ReduceOp.MAX)
        dist.all_reduce(finite, op=dist.ReduceOp.MIN)
    sse, ssr = (float(v) for v in sums)
    max_err, max_ref = (float(v) for v in maxs)
    numel = world * layers * 2 * eng.n * eng.c
    mse = sse / numel
    ok = bool(finite.item()) and all(math.isfinite(v) for v in (sse, ssr, max_err, max_ref))
    return {"tensors": f"every layer's K and V ({2 * layers} tensors per rank, all {world} ranks): the reconstruction after the "
                       "last timed step against its raw input",
            "rel_l2": math.sqrt(sse / ssr) if ok and ssr > 0 else None,
            "rel_l2_worst_tensor": float(worst_rel) if ok else None,
            "max_abs": max_err if ok else None,
            "psnr_db": (10.0 * math.log10(max_ref * max_ref / mse)) if ok and mse > 0 else None,
            "finite": ok}


def tensor_hash(t):
    """64-bit position-weighted wrap-around checksum of an fp16 tensor's bytes (plumbing: plain torch integer ops)."""
    w = t.reshape(-1).view(torch.int64)
    idx = torch.arange(1, 2 * w.numel(), 2, dtype=torch.int64, device=t.device)  # odd multipliers
    return (w * idx).sum()


def ranks_identical(eng, world, device):
    """The error-feedback invariant (main.py:398-419): after a step EVERY rank holds bit-identical reconstructions
    of every origin's shard.  All-gather one 64-bit checksum per global buffer and compare."""
    h = torch.stack([tensor_hash(g) for g in eng.global_k + eng.global_v])
    if world == 1:
        return True, int(h.numel())
    allh = [torch.empty_like(h) for _ in range(world)]
    dist.all_gather(allh, h)
    return all(bool(torch.equal(a, allh[0])) for a in allh), int(h.numel())


def oracle_parity(eng, xk, xv, ctype, codec, world, rank, device, barrier):
    """One more step of layer 0, outside the timed region, checked by the CPU oracle on rank 0 (bench.py may run
    `oracle/` as the checker): the payloads of ALL origins as they arrived in rank 0's receive memory -- over
    NVLink for the peers -- must carry exactly the sign bits / codes the reference computes from (x - base),
    scales within 1 fp16 ulp, and rank 0's reconstruction of every origin must be bit-identical to the
    reference's dequant of that payload against the base cached before the step (oracle/check.py)."""
    n, c = eng.n, eng.c
    barrier()
    base_k, base_v = eng.global_k[0].clone(), eng.global_v[0].clone()
    eng.exchange(0, xk, xv, ctype)
    barrier()
    payloads = [[eng.slot_bytes(0, r, ctype, j).clone() for j in range(2)] for r in range(world)]
    xs = [xk.reshape(n, c), xv.reshape(n, c)]
    if world > 1:  # the peers' raw shards, for the sender-side checks (plumbing: NCCL all-gather)
        gathered = []
        for x in xs:
            buf = torch.empty((world * n, c), dtype=torch.half, device=device)
            dist.all_gather_into_tensor(buf, x.contiguous())
            gathered.append(buf)
    else:
        gathered = xs
    if rank != 0:
        return None
    from oracle import check as ocheck
    out = {}
    for j, (name, base, glob) in enumerate((("k", base_k, eng.global_k[0]), ("v", base_v, eng.global_v[0]))):
        sh = lambda t, r: t[r * n:(r + 1) * n].cpu()  # noqa: E731
        out[name] = ocheck.check_exchange(codec, [sh(gathered[j], r) for r in range(world)],
                                          [sh(base, r) for r in range(world)],
                                          [payloads[r][j].cpu().numpy() for r in range(world)],
                                          [sh(glob, r) for r in range(world)])
    return {"ok": out["k"]["ok"] and out["v"]["ok"], "layer": 0, "rows_checked": 2 * world * n,
            "what": "rank 0: payloads of all origins as received + reconstructions of layer 0 vs oracle/check.py "
                    "(codes bit-exact, V <= 1 ulp and U <= 2 ulp, reconstruction bit-exact)", **out}


def measure_kernels(args, eng, ks, vs, ctype, world, rank, n_local, layers, transport, barrier):
    """Per-kernel durations and the roofline of the dominant one.

    Every kernel of the step is timed on its own: a CUDA graph holding that kernel's launch for ALL
    layers (distinct buffers per layer: `layers` x tens of MB >> 126 MB L2, so every launch is cold) is
    replayed between two CUDA events on the launching stream.  No per-launch events (they add a front-end
    round trip of several us to a 10-20 us kernel)."""
    versions = len(ks)
    from compactfusion_b200 import _native as nv
    e_tensor = n_local * CH
    per_byte = 8 if args.codec == "binary" else 4
    vsel = args.steps % versions
    part_b = 148 // 2  # column/token partials written by pass 1 (one row block per CTA)
    kernels = []

    def time_kernel(name, fn, algo_bytes, reps=3):
        barrier()
        for layer in range(layers):
            fn(layer)
        torch.cuda.synchronize()
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for layer in range(layers):
                    fn(layer)
            run = g.replay
        except Exception:
            torch.cuda.synchronize()

            def run():
                for layer in range(layers):
                    fn(layer)
        run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            run()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / (reps * layers)
        kernels.append({"kernel": name, "avg_launch_us": us, "algorithmic_bytes_per_launch": algo_bytes,
                        "achieved": algo_bytes / us / 1e3})

    stats_name = "k_delta_stats_tma" if os.environ.get("CF_LEGACY_KERNELS", "0") != "1" else "k_delta_stats"
    apply_name = "k_apply_codes_tma" if os.environ.get("CF_LEGACY_KERNELS", "0") != "1" else "k_apply_codes"
    fused = world > 1 and eng.fused(ctype)
    comp = eng.compress_put if fused else eng.compress
    fan = world if fused else 1  # fused put: codes and scales are stored to all W receive slots (W-1 over NVLink)
    # pass 1 over K and V of this rank: read x and base, write sign bits (BINARY) + partial sums
    time_kernel(stats_name, lambda l: comp(l, ks[vsel][l], vs[vsel][l], ctype, nv.PASS_STATS),
                2 * (4 * e_tensor + (fan * e_tensor // 8 if args.codec == "binary" else 0) + 2 * n_local + 4 * part_b * CH))
    time_kernel("k_finalize_scales", lambda l: comp(l, ks[vsel][l], vs[vsel][l], ctype, nv.PASS_FINALIZE),
                2 * (4 * part_b * CH + 2 * n_local + fan * 2 * (n_local + CH)))
    if fused:
        for k_ in kernels:
            k_["fused_put"] = True
            k_["nvlink_bytes_per_launch"] = (world - 1) * 2 * (
                (e_tensor // 8 if args.codec == "binary" else 0) if k_["kernel"] == stats_name else 2 * (n_local + CH))
    if args.codec == "int2":
        time_kernel("k_int2_encode_tma", lambda l: comp(l, ks[vsel][l], vs[vsel][l], ctype, nv.PASS_ENCODE),
                    2 * (4 * e_tensor + fan * e_tensor // 4 + 2 * (n_local + CH)))
        if fused:
            kernels[-1]["fused_put"] = True
            kernels[-1]["nvlink_bytes_per_launch"] = (world - 1) * 2 * (e_tensor // 4)
    if transport == "p2p" and not fused:
        # one-sided exchange: this rank's [K payload | V payload] stored into all W receive slots; W-1 of them
        # cross NVLink (measured peer-copy peak 770 GB/s per direction, B200_PROFILING.md)
        slot_bytes = 2 * (e_tensor // per_byte + 2 * (n_local + CH))
        time_kernel("k_p2p_put", lambda l: eng.gather(ctype, l), (world + 1) * slot_bytes)
        kernels[-1]["nvlink_bytes_per_launch"] = (world - 1) * slot_bytes
        kernels[-1]["nvlink_gbs"] = (world - 1) * slot_bytes / kernels[-1]["avg_launch_us"] / 1e3
        kernels[-1]["nvlink_frac_of_770"] = kernels[-1]["nvlink_gbs"] / 770.0
    # reconstruct K and V of all W origins in place: read base + codes + scales, write recon
    if MODE == "ring":
        n_launch_per_call = world

        def dec(l):
            for r in range(world):
                eng.decompress(l, ctype, origins=(eng.hop_origin(r),))
    else:
        n_launch_per_call = (2 * world + 15) // 16

        def dec(l):
            eng.decompress(l, ctype)
    time_kernel(apply_name, dec,
                2 * world * (2 * e_tensor + e_tensor // per_byte + 2 * (n_local + CH) + 2 * e_tensor))
    note(rank, "per-kernel timing done")
    kernels[-1]["avg_launch_us"] /= n_launch_per_call
    kernels[-1]["algorithmic_bytes_per_launch"] //= n_launch_per_call
    peak, peak_src = measured_hbm_peak()
    k_total = sum(k["avg_launch_us"] * (n_launch_per_call if k["kernel"] == apply_name else 1) for k in kernels)
    for k in kernels:
        mult = n_launch_per_call if k["kernel"] == apply_name else 1
        k["frac"] = k["achieved"] / peak
        k["share_of_kernel_time"] = k["avg_launch_us"] * mult / k_total
    dom = max(kernels, key=lambda k: k["share_of_kernel_time"])
    roofline = {"bound": "hbm", "achieved": dom["achieved"], "peak": peak, "unit": "GB/s", "frac": dom["frac"],
                "traffic": ncu_traffic(dom["kernel"], args.codec, n_local, world), "kernel": dom["kernel"],
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                "avg_launch_us": dom["avg_launch_us"], "share_of_step": dom["share_of_kernel_time"],
                "peak_source": peak_src, "kernels": kernels,
                "method": "each kernel alone: one CUDA graph with its launch for all layers (cold buffers), "
                          "replayed 3x between two CUDA events"}
    return roofline


def measure_e2e(args, eng, sample, ctype, world, n_local, layers, device, transport, barrier):
    """The same step with K/V in pinned host memory: H2D -> exchange -> D2H of the reconstructed global K/V,
    double-buffered over three streams (copy in, compute, copy out).  Wall clock, max over ranks."""
    e2e_layers = layers
    hk = [torch.empty((n_local, CH), dtype=torch.half).pin_memory() for _ in range(2)]
    hv = [torch.empty((n_local, CH), dtype=torch.half).pin_memory() for _ in range(2)]
    for b_ in hk + hv:
        b_.copy_(sample.cpu())
    out_k = torch.empty((world * n_local, CH), dtype=torch.half).pin_memory()
    out_v = torch.empty((world * n_local, CH), dtype=torch.half).pin_memory()
    dk = [torch.empty((n_local, CH), dtype=torch.half, device=device) for _ in range(2)]
    dv = [torch.empty((n_local, CH), dtype=torch.half, device=device) for _ in range(2)]
    copy_in, copy_out = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
    main_s = torch.cuda.current_stream()
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        in_done = [None, None]
        dones = []
        for layer in range(e2e_layers):
            s = layer & 1
            if layer >= 2:
                copy_in.wait_event(dones[layer - 2])  # staging buffer s was last read by the exchange of layer - 2
            with torch.cuda.stream(copy_in):
                dk[s].copy_(hk[s], non_blocking=True)
                dv[s].copy_(hv[s], non_blocking=True)
                in_done[s] = torch.cuda.Event()
                in_done[s].record(copy_in)
            main_s.wait_event(in_done[s])
            gk, gv = eng.exchange(layer, dk[s], dv[s], ctype)
            done = torch.cuda.Event()
            done.record(main_s)
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(done)
                out_k.copy_(gk, non_blocking=True)
                out_v.copy_(gv, non_blocking=True)
            dones.append(done)
        main_s.wait_stream(copy_out)
        copy_in.wait_event(dones[-1])
        if len(dones) > 1:
            copy_in.wait_event(dones[-2])

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": job_bytes(layers, world) / e2e_s / 1e9, "unit": UNIT,
           "h2d_bytes_per_step": world * layers * 2 * n_local * CH * 2,
           "d2h_bytes_per_step": world * layers * 2 * world * n_local * CH * 2,
           "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
           "path": f"pinned host K/V -> H2D -> {type(eng).__name__}.exchange (C-ABI batched kernels, transport "
                   f"{transport}) -> D2H of reconstructed global K/V, double-buffered over 3 streams"}
    return e2e


def gpu_reference(args, ks, vs, pattern, layers, n_local, device, steps=3):
    """The kernel-to-beat, in the same run on the same GPU: the UNMODIFIED reference (baseline/_ref: Triton fastpath
    kernels + eager torch scale passes, its own compact_compress / compact_decompress, per-call allocation and
    Python dispatch as shipped) doing this workload's single-GPU step -- per layer, K and V: compress without cache
    update, decompress with cache update (compact_all_gather at world size 1, main.py:390-420).  CUDA events around
    `steps` full steps after one untimed step (Triton JIT + warm-up).  None if baseline/_ref is not staged."""
    ref = _load_reference_main()
    if ref is None:
        return None
    cm, RT, RCfg = ref
    ct = RT(args.codec)
    cm.compact_init(RCfg(enabled=True, residual=1, ef=True, simulate=False, fastpath=True, comp_rank=-1,
                         compress_func=lambda l, s: ct))
    shape = (1, n_local, 1, CH)

    def step(v):
        for l in range(layers):
            for tag, x in ((f"{l}-k", ks[v][l]), (f"{l}-v", vs[v][l])):
                p = cm.compact_compress(f"{tag}-0", x.view(shape), ct, update_cache=False)
                cm.compact_decompress(f"{tag}-0", p, ct, shape, update_cache=True)

    for l in range(layers):  # the reference's WARMUP step: cache the raw tensors
        cm.compact_decompress(f"{l}-k-0", ks[0][l].view(shape), RT.WARMUP, shape, update_cache=True)
        cm.compact_decompress(f"{l}-v-0", vs[0][l].view(shape), RT.WARMUP, shape, update_cache=True)
    step(1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        step(pattern[(i + 2) % len(pattern)])
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    cm.compact_reset()
    return {"ms_per_step": ms, "value": job_bytes(layers, 1) / (ms * 1e-3) / 1e9, "unit": UNIT, "steps": steps,
            "what": "unmodified reference (baseline/_ref: Triton fastpath + eager torch, compact_compress / "
                    "compact_decompress as shipped) on the same GPU, same inputs, world size 1"}


class RawBaseline:
    """`--codec raw`: what xDiT moves WITHOUT the plugin, same K / V, same harness.  None of
    our kernels run; the lines say `"impl": "uncompressed_baseline"`.
      allgather : synchronous all-gather of the fp16 shards, K and V per layer (patchpara/fwd.py:103-111)
      ring      : the uncompressed ring -- every layer relays the raw [K | V] block W-1 hops with batch_isend_irecv
                  (xfuser/core/long_ctx_attention/ring/ring_flash_attn.py:16-137, attention left out)
      async     : DistriFusion's stale all-gather -- the collective of step t is waited for in step t+1, the
                  attention of step t uses the peers' previous K / V (patchpara/fwd.py:113-173)"""

    def __init__(self, kind, eng, world, rank, layers, n_local, device):
        self.kind, self.eng, self.world, self.rank, self.layers = kind, eng, world, rank, layers
        self.kernel_launches = 0
        if kind == "ring" and world > 1:
            self.buf = [torch.empty((2, n_local, CH), dtype=torch.half, device=device) for _ in range(2)]
        if kind == "async":
            self.pending = [None] * layers
            self.nxt_k = [torch.empty_like(g) for g in eng.global_k]
            self.nxt_v = [torch.empty_like(g) for g in eng.global_v]

    def step(self, ks, vs, _ctype=None, _overlap=False):
        eng, W = self.eng, self.world
        for l in range(self.layers):
            k2, v2 = ks[l].reshape(eng.n, eng.c), vs[l].reshape(eng.n, eng.c)
            if W == 1 or self.kind == "allgather":
                eng.warmup(l, k2, v2)
            elif self.kind == "ring":
                cur = self.buf[0]
                cur[0].copy_(k2)
                cur[1].copy_(v2)
                send_to, recv_from = (self.rank + 1) % W, (self.rank - 1) % W
                for s in range(W):
                    origin = (self.rank - s) % W
                    eng._shard(eng.global_k[l], origin).copy_(cur[0])  # hand the block to "attention"
                    eng._shard(eng.global_v[l], origin).copy_(cur[1])
                    if s + 1 < W:
                        nxt = self.buf[(s + 1) & 1]
                        for r in dist.batch_isend_irecv([dist.P2POp(dist.isend, cur, send_to),
                                                         dist.P2POp(dist.irecv, nxt, recv_from)]):
                            r.wait()
                        cur = nxt
            else:  # stale-async
                if self.pending[l] is not None:
                    for h in self.pending[l]:
                        h.wait()
                    eng.global_k[l], self.nxt_k[l] = self.nxt_k[l], eng.global_k[l]   # the peers' previous step
                    eng.global_v[l], self.nxt_v[l] = self.nxt_v[l], eng.global_v[l]
                eng._shard(eng.global_k[l], self.rank).copy_(k2)                       # fresh local shard
                eng._shard(eng.global_v[l], self.rank).copy_(v2)
                self.pending[l] = [dist.all_gather_into_tensor(self.nxt_k[l], k2, async_op=True),
                                   dist.all_gather_into_tensor(self.nxt_v[l], v2, async_op=True)]

    def finish(self):
        if self.kind == "async":
            for hs in self.pending:
                for h in hs or []:
                    h.wait()
            self.pending = [None] * self.layers


class DropinDriver:
    """`--api dropin`: the step as an xDiT pipeline drives it -- `compact_fwd(q, k, v, ..., mod_idx, current_iter)` once
    per attention layer (ring.py:36-70 -> _compact_ring_fwd / patch_gather_fwd), eager launches on the current
    stream, the plugin's own state machine deciding WARMUP vs codec from `compress_func(layer, step)`.  The attention
    block behind the exchange is replaced by a stub that launches nothing (`attention.set_attention_override`), so
    the figure is the exchange hooks' cost, comparable with the engine lines."""

    def __init__(self, ctype, layers, n_local, device, transport):
        import compactfusion_b200 as cf
        from compactfusion_b200 import attention, dropin
        from compactfusion_b200.utils import COMPACT_COMPRESS_TYPE as T
        w = WORKLOADS[WORKLOAD]
        self.cf, self.dropin, self.layers = cf, dropin, layers
        self.bs, self.h = w["bs"], w["heads"]
        self.shape = (self.bs, n_local // self.bs, self.h, CH // self.h)
        os.environ["CF_DROPIN_TRANSPORT"] = transport
        patch = MODE == "patch"
        cfg = cf.CompactConfig(enabled=True, override_with_patch_gather_fwd=patch,
                               patch_gather_fwd_config=cf.PatchConfig(True, False, 1) if patch else None,
                               compress_func=lambda l, s: ctype if s >= 1 else T.WARMUP, comp_rank=-1, residual=1,
                               ef=True, fastpath=True)
        cf.compact_init(cfg)
        # like a production run of the reference: its scope profiler records two CUDA events per hook call unless
        # switched off (xfuser/prof.py:14-16)
        from compactfusion_b200.prof import Profiler
        Profiler.instance().disable()
        self.q = torch.zeros(self.shape, dtype=torch.half, device=device)
        out = torch.zeros(self.shape, dtype=torch.half, device=device)
        lse = torch.zeros((self.bs, self.h, self.shape[1]), dtype=torch.float32, device=device)
        attention.set_attention_override(lambda q, k, v, *a: (out, lse))
        self.it = 0
        self._views = {}
        self.kernel_launches_base = 0

    @property
    def eng(self):
        return self.dropin.engines()[0]

    def step(self, ks, vs, _ctype=None, _overlap=False):
        """One denoising step: every layer's hook call.  The step index decides WARMUP (step 0) vs the codec."""
        self.cf.compact_set_step(self.it)
        kv = self._views.get(id(ks))
        if kv is None:   # a model hands over (bs, s, h, d) tensors: build the views of this input version once
            kv = self._views[id(ks)] = ([k.view(self.shape) for k in ks[:self.layers]], [v.view(self.shape) for v in vs[:self.layers]], ks, vs)
        k4, v4, fwd, q, it = kv[0], kv[1], self.cf.compact_fwd, self.q, self.it
        for l in range(self.layers):
            fwd(q, k4[l], v4[l], causal=False, mod_idx=l, current_iter=it)
        self.it += 1


def dropin_requested(args):
    return args.api == "dropin"


def measure_lowrank(args, eng, ks, vs, ctype, world, layers, barrier):
    """Low-rank codecs: the two phases of a layer timed over all layers (eager launches, CUDA events): project
    (+ int4 of the factors, + put) and reconstruct.  Algorithmic bytes of the projector: delta must be streamed 4
    times at least (2 iterations, U, V: SURVEY.md section 8d) from x and base, 4 x 4E, plus the payload."""
    e = eng.n * CH
    vsel = 2

    def timed(fn, reps=2):
        barrier()
        for l in range(layers):
            fn(l)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            for l in range(layers):
                fn(l)
        b.record()
        barrier()
        return a.elapsed_time(b) * 1e3 / (reps * layers)

    payload = eng._numel(ctype) * 2
    us_c = timed(lambda l: eng.send(l, ks[vsel][l], vs[vsel][l], ctype))
    us_d = timed(lambda l: eng.decompress(l, ctype))
    kernels = [
        {"kernel": "cf_lowrank_project (+ int4 of U, V^T) x {K, V}", "avg_launch_us": us_c,
         "algorithmic_bytes_per_launch": 2 * (16 * e + payload), "achieved": 2 * (16 * e + payload) / us_c / 1e3},
        {"kernel": "cf_lowrank_reconstruct x {K, V} x origins", "avg_launch_us": us_d,
         "algorithmic_bytes_per_launch": 2 * world * (4 * e + payload), "achieved": 2 * world * (4 * e + payload) / us_d / 1e3},
    ]
    peak, peak_src = measured_hbm_peak()
    tot = us_c + us_d
    for k in kernels:
        k["frac"] = k["achieved"] / peak
        k["share_of_kernel_time"] = k["avg_launch_us"] / tot
    dom = max(kernels, key=lambda k: k["share_of_kernel_time"])
    return {"bound": "hbm", "achieved": dom["achieved"], "peak": peak, "unit": "GB/s", "frac": dom["frac"], "traffic": None,
            "kernel": dom["kernel"], "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
            "avg_launch_us": dom["avg_launch_us"], "share_of_step": dom["share_of_kernel_time"], "peak_source": peak_src,
            "kernels": kernels, "method": "each phase of a layer (several launches) over all layers, eager, CUDA events"}


def run_config1(args, device):
    """BASELINE configs[0] on one GPU: `compact_compress` (INT4, residual 1, error feedback, cache update) +
    `compact_decompress` on the receiver's key, per step, through the plugin API (`main.py:169,322`); x_t = x_0 +
    0.05 t randn, step 0 is WARMUP.  One "step" of the JSON line = one denoising step of this
    one-tensor workload; the timed region is steps 1..27 of a 28-step series, repeated `--steps` times."""
    import compactfusion_b200 as cf
    from compactfusion_b200.quality import error_stats
    T = cf.COMPACT_COMPRESS_TYPE
    n, c, series = SEQ, CH, 28
    g = torch.Generator(device=device).manual_seed(0)
    x0 = torch.randn(n, c, generator=g, device=device)
    xs = [(x0 + 0.05 * t * torch.randn(n, c, generator=g, device=device)).half().view(1, n, 24, c // 24) for t in range(series)]
    cf.compact_init(cf.CompactConfig(enabled=True, residual=1, ef=True, simulate=False, comp_rank=-1,
                                     compress_func=lambda l, s: T.INT4 if s >= 1 else T.WARMUP))
    from compactfusion_b200.prof import Profiler
    Profiler.instance().disable()   # as in a production run of the reference (xfuser/prof.py:14-16): no CUDA events per scope

    def one_series():
        for t in range(series):
            ct = T.INT4 if t >= 1 else T.WARMUP
            p = cf.compact_compress("0-0-k", xs[t], ct, update_cache=True)
            rec = cf.compact_decompress("1-0-k", p, ct, xs[t].shape, update_cache=True)
        return rec

    for _ in range(max(args.warmup, 3)):
        one_series()
    clocks = ClockSampler(0)
    clocks.start()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        rec = one_series()
    b.record()
    torch.cuda.synchronize()
    ms_step = a.elapsed_time(b) / (args.steps * (series - 1))   # per compressed step (27 of the 28 are compressed)
    clock_info = clocks.stop()
    fid = error_stats(rec.reshape(n, c), xs[-1].reshape(n, c))
    same = bool(torch.equal(cf.compact_cache().get_base("0-0-k").reshape(n, c), cf.compact_cache().get_base("1-0-k").reshape(n, c)))
    e = n * c
    algo = (6 * e + e // 2 + 4 * c) + (4 * e + e // 2 + 4 * c)   # compress+EF, decompress+update
    peak, peak_src = measured_hbm_peak()
    cpu = None
    if not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline("int4", 1, 1, 1, budget_s=8.0)
        except Exception as ex:  # noqa: BLE001
            cpu = {"error": f"{type(ex).__name__}: {ex}"[:200]}
    print(json.dumps({
        "metric": METRIC, "value": 2 * e / (ms_step * 1e-3) / 1e9, "unit": UNIT, "n_gpus": 1, "steps": args.steps * (series - 1),
        "warmup": max(args.warmup, 3) * series, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "api": "plugin (compact_compress / compact_decompress, eager launches)", "codec": "int4",
                   "layers": 1, "seq": n, "channels": c, "world": 1, "series_steps": series, "launch_mode": "eager",
                   "l2": "one 25 MB tensor + its two cached bases per step: L2-resident by construction of configs[0]"},
        "roofline": {"bound": "hbm", "achieved": algo / (ms_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": algo / (ms_step * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "kernel": "cf_int4_compress (+EF) + cf_int4_decompress: the step's 5 kernels + payload cat",
                     "algorithmic_bytes_per_launch": algo, "avg_launch_us": ms_step * 1e3},
        "cpu_baseline": cpu, "e2e": None, "gpu_launches": args.steps * (series - 1) * 5, "clocks": clock_info,
        "fidelity": {"rel_l2": fid["rel_l2"], "max_abs": fid["max_abs"], "psnr_db": fid["psnr_db"], "finite": True},
        "parity_ok": bool(same and fid["rel_l2"] < 0.05),
        "ranks_identical": {"ok": same, "what": "sender cache == receiver cache after 28 steps (bit-exact)"},
    }))


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.hang_dump > 0:
        import faulthandler
        faulthandler.dump_traceback_later(args.hang_dump, exit=True)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.layers = select_workload(args.workload, args.layers)
    if MODE == "roundtrip":
        args.codec = "int4"   # configs[0] names its codec
    assert args.codec != "int4" or MODE == "roundtrip", "--codec int4 is the config1_int4_roundtrip workload's codec"
    if args.impl == "reference":
        run_reference(args, max(world, args.gpus), rank)
        return
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun for N > 1"
    assert SEQ % world == 0
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product has no CPU path)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if MODE == "roundtrip":
        assert world == 1, "configs[0] is a world-size-1 workload"
        from compactfusion_b200 import build as cf_build
        if not os.path.exists(cf_build.OUT):
            cf_build.build()
        run_config1(args, device)
        return
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    elif args.api == "dropin":
        # the reference's hooks ask torch.distributed for rank / world size (patchpara/fwd.py:60-61): a one-rank group
        import socket
        with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sock:
            sock.bind(("127.0.0.1", 0))
            port = sock.getsockname()[1]
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", world_size=1, rank=0)

    from compactfusion_b200 import build as cf_build
    if not os.path.exists(cf_build.OUT) and local_rank == 0:
        cf_build.build()  # a checkout without the build artefact: compile it (no fallback: failure is fatal)
    if world > 1:
        dist.barrier()
    from compactfusion_b200.engine import PatchGatherEngine, RingExchangeEngine
    from compactfusion_b200.utils import COMPACT_COMPRESS_TYPE as T
    raw = args.codec == "raw"
    lowrank = args.codec.startswith("lowrank")
    comp_rank = int(args.codec.lstrip("lowrankq")) if lowrank else None
    ctype = (T.LOW_RANK_Q if args.codec.startswith("lowrankq") else T.LOW_RANK) if lowrank else \
        {"binary": T.BINARY, "int2": T.INT2, "raw": T.WARMUP}[args.codec]
    if lowrank:
        # the projector draws its random start and allocates per call (like subspace_iter, compress_lowrank.py:41):
        # eager launches; the oracle leg is bit-exactness of the sign codecs, not applicable to a random subspace
        args.no_graph, args.no_parity = True, True
        assert not dropin_requested(args), "--api dropin with a low-rank codec: use the engine line"
    n_local, layers = SEQ // world, args.layers
    # ring workloads consume the origins hop by hop (one flag-waiting decompress launch per origin);
    # patch workloads reconstruct all origins in one launch
    engine_cls = RingExchangeEngine if MODE == "ring" else PatchGatherEngine
    probe_note = None
    dropin_api = args.api == "dropin"
    assert not (dropin_api and raw), "--api dropin drives the compressed hooks; use --api engine for --codec raw"
    if world > 1 and args.transport == "auto" and not raw and not dropin_api:
        ok, why = p2p_probe(engine_cls, n_local, ctype, device, comp_rank)
        if not ok:
            args.transport, probe_note = "nccl", "one-sided transport rejected by the probe step: " + (why or "a peer failed")
        note(rank, f"p2p probe: {'ok' if ok else probe_note}")
    # inputs_stable: the bench's K/V inputs are static buffers, so the early pipeline fill is legitimate here
    driver = None
    if dropin_api:
        driver = DropinDriver(ctype, layers, n_local, device, args.transport)
        args.no_graph, eng, transport = True, None, "pending"
    else:
        eng = engine_cls(layers, n_local, CH, group=None, device=device, transport=args.transport, inputs_stable=True,
                         comp_rank=comp_rank)
        transport = eng.prepare(ctype) if world > 1 else "none (single GPU)"
    if raw and world > 1:
        transport = "nccl"  # all_gather_into_tensor of the raw fp16 shards (engine.warmup)
    note(rank, f"transport: {transport}")
    if transport == "nccl" and not args.no_graph:
        # NCCL collectives inside the captured step hang on replay on this stack (torch 2.11 / NCCL 2.28):
        # with the NCCL transport the step is launched eagerly
        args.no_graph = True
    # 4 consecutive AR(1) time steps of every K / V, walked back and forth (0 1 2 3 2 1 0 1 ...): every step
    # compresses a tensor one AR step away from what the cache was built from (a stationary Gaussian AR(1)
    # process is time-reversible), so the residual statistics of a real denoising loop hold at every step
    versions = 4
    pattern = list(range(versions)) + list(range(versions - 2, 0, -1))
    acts = synth_activations(n_local, layers, versions, device, rank)
    ks = [[acts[l][0][v] for l in range(layers)] for v in range(versions)]
    vs = [[acts[l][1][v] for l in range(layers)] for v in range(versions)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    note(rank, "inputs ready")
    # step 0: WARMUP (uncompressed), not timed
    if dropin_api:
        driver.step(ks[0], vs[0])
        eng = driver.eng  # the engine the hooks built while the warm-up step walked the layers
    else:
        eng.step(ks[0], vs[0], T.WARMUP)
    stepper = driver if dropin_api else eng
    if raw:
        stepper = RawBaseline(args.raw_exchange, eng, world, rank, layers, n_local, device)
    torch.cuda.synchronize()
    note(rank, "warmup step done")
    # first compressed step, eagerly, on the NEXT version: loads the kernels and sizes the workspaces before any
    # capture.  (Never compress a tensor against an identical base: delta == 0 gives the reference's 0/0 BINARY
    # token scale, fastpath.py:164-165, and the NaN would stay in the error-feedback cache.)
    if not raw:
        stepper.step(ks[1], vs[1], ctype, args.overlap)
        torch.cuda.synchronize()
    if dropin_api:
        transport = eng.transport if world > 1 else "none (single GPU)"
    graphs = None
    mode = "eager (compact_fwd hook per layer)" if dropin_api else "eager"
    if not args.no_graph:
        try:
            graphs = [eng.capture_step(ks[v], vs[v], ctype, warmup_iters=0, overlap=args.overlap)
                      for v in range(versions)]
            mode = "cuda_graph"
        except Exception as e:  # capture can fail with NCCL inside: fall back to eager launches
            graphs, mode = None, f"eager (graph capture failed: {type(e).__name__})"
            torch.cuda.synchronize()

    note(rank, f"launch mode: {mode}")

    step_no = [0]

    def run_step(_i=None):
        step_no[0] += 1
        v = pattern[(step_no[0] + 1) % len(pattern)]  # 2 3 2 1 0 1 2 ... after the eager step on version 1
        run_step.last = v
        if graphs is not None:
            graphs[v].replay()
        else:
            stepper.step(ks[v], vs[v], ctype, args.overlap)

    if raw:
        mode = f"eager (uncompressed {args.raw_exchange})"
    n_warm = max(args.warmup, 3)
    if dropin_api:
        # the hooks capture one small CUDA graph per (layer, K/V address pair) at the second sighting of a pair
        # (engine._graphed): walk the whole version pattern twice before the timed region, so that it measures the
        # steady state of a denoising loop and not the one-time captures
        n_warm = max(n_warm, 2 * len(pattern) + 1)
    for i in range(n_warm):
        run_step(i)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    eng.kernel_launches = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    host_t0 = time.perf_counter()
    host_first = None
    for i in range(args.steps):
        run_step(i)
        if i == 1:   # two steps fit the launch queue: how long the HOST needs to enqueue a step (eager modes)
            host_first = (time.perf_counter() - host_t0) / 2
    if raw:
        stepper.finish()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    note(rank, f"timed region done: {ms / args.steps:.3f} ms/step"
         + (f"; host enqueue of a step {host_first * 1e3:.3f} ms" if host_first else ""))
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clock_info = clocks.stop() if rank == 0 else None
    if getattr(eng, "plan_host_s", None) and eng.plan_host_s[1]:
        note(rank, f"CF_PLAN_TIMING: {eng.plan_host_s[0] / eng.plan_host_s[1] * 1e6:.1f} us of host time per layer inside the C calls")
    launches = (eng.launches_per_graph * args.steps) if graphs is not None else eng.kernel_launches
    ms_per_step = ms / args.steps
    value = job_bytes(layers, world) / (ms_per_step * 1e-3) / 1e9

    v_last = run_step.last
    fidelity = measure_fidelity(eng, ks[v_last], vs[v_last], rank, world, device)
    identical, n_hashed = ranks_identical(eng, world, device)
    parity = None
    if not raw and not args.no_parity:
        v_next = pattern[(step_no[0] + 2) % len(pattern)]
        try:
            parity = oracle_parity(eng, ks[v_next][0], vs[v_next][0], ctype, args.codec, world, rank, device, barrier)
        except Exception as e:  # noqa: BLE001 -- reported, and parity_ok goes false
            parity = {"ok": False, "error": f"{type(e).__name__}: {e}"}
    note(rank, "fidelity / identity / oracle parity done")
    if raw:
        roofline = None
    elif lowrank:
        roofline = measure_lowrank(args, eng, ks, vs, ctype, world, layers, barrier)
    else:
        roofline = measure_kernels(args, eng, ks, vs, ctype, world, rank, n_local, layers, transport, barrier)
    e2e = None if args.no_e2e else measure_e2e(args, eng, acts[0][0][0], ctype, world, n_local, layers, device,
                                               transport, barrier)
    note(rank, "e2e done")
    gpu_ref = None
    if world == 1 and not raw and not lowrank and not args.no_gpu_reference:
        try:
            gpu_ref = gpu_reference(args, ks, vs, pattern, layers, n_local, device)
        except Exception as e:  # noqa: BLE001 -- a side figure: report why it is missing
            gpu_ref = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.synchronize()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not raw:
        try:
            cpu = cpu_baseline(args.codec, world, layers, args.cpu_sample_layers)
        except Exception as e:
            cpu = {"error": f"{type(e).__name__}: {e}"}

    timeouts = bool(eng.p2p_error())
    if world > 1:
        t = torch.tensor([int(timeouts)], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        timeouts = bool(t.item())
    if rank == 0:
        print(json.dumps({
            **({"impl": "uncompressed_baseline"} if raw else {}),
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "api": ("dropin (compact_fwd hooks, eager launches, attention stubbed out)"
                                                     if dropin_api else "engine (whole-step runtime)"),
                       "exchange": MODE, "codec": args.codec, "layers": layers, "seq": SEQ,
                       "channels": CH, "world": world, "shard_rows": n_local, "launch_mode": mode,
                       "host_enqueue_ms_per_step": (round(host_first * 1e3, 4) if (host_first and graphs is None) else None),
                       "schedule": "two chains (compress | reconstruct)" if (args.overlap and not raw and eng.can_overlap(ctype)) else "serial",
                       "transport": transport + (" (fused into the codec kernels)" if world > 1 and eng.fused(ctype) else ""),
                       **({"transport_note": probe_note} if probe_note else {}),
                       "l2": f"inputs larger than L2 (each step touches {(1 + world) * layers * 2 * n_local * CH * 2 / 1e9:.1f} GB "
                             "of distinct K/V inputs + cached bases per rank)"},
            "roofline": roofline, "cpu_baseline": cpu, "gpu_reference": gpu_ref, "e2e": e2e, "gpu_launches": launches,
            "clocks": clock_info,
            "fidelity": fidelity,
            "ranks_identical": {"ok": identical, "buffers_hashed": n_hashed,
                                "what": "64-bit checksum of every layer's global K and V buffer, equal on all ranks"},
            "oracle_parity": parity,
            "parity_ok": bool(fidelity.get("finite") and identical and not timeouts
                              and (parity is None or parity.get("ok"))),
            "p2p_wait_timeouts": timeouts,
        }))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
