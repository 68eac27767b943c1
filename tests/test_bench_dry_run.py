"""Dry run of bench.py's GPU arm on the CPU: CUDA streams / events / graphs, the native library and (for the
N = 2 case) torch.distributed are replaced by recording stand-ins (TEST ONLY: nothing is computed, times are
made up), so that every line of the host code that produces the round's JSON line executes here -- argument
handling, the transport probe, engine setup, graph capture, per-kernel timing legs, fidelity, e2e staging,
JSON assembly -- and a typo cannot first show up on the GPU box."""
import contextlib
import ctypes
import json
import sys
import types

import pytest
import torch

from test_engine_schedule_fake_cuda import FakeCuda, FakeEvent, FakeLib, FakeStream


class TimedEvent(FakeEvent):
    def __init__(self, enable_timing=False, **k):
        super().__init__()

    def elapsed_time(self, other):
        return 1.0  # ms


class FakeGraph:
    def replay(self):
        return None


@contextlib.contextmanager
def fake_graph_ctx(g, *a, **k):
    yield


class IpcLib(FakeLib):
    """FakeLib whose cf_ipc_alloc / cf_ipc_open hand out made-up pointers."""
    next_ptr = 0x7000000000

    def __getattr__(self, name):
        if name in ("cf_ipc_alloc", "cf_ipc_open"):
            def alloc(*args):
                out = args[1]
                IpcLib.next_ptr += 0x100000000
                ctypes.cast(out, ctypes.POINTER(ctypes.c_void_p))[0] = IpcLib.next_ptr
                return 0
            return alloc
        if name == "cf_last_error":
            return lambda: b""
        return super().__getattr__(name)


class FakeDist:
    """torch.distributed for ONE process that believes it is rank 0 of `world`."""

    class ReduceOp:
        MAX, MIN, SUM = "max", "min", "sum"

    def __init__(self, world):
        self.world, self.up = world, False

    def init_process_group(self, *a, **k):
        self.up = True

    def is_initialized(self):
        return self.up

    def get_world_size(self, group=None):
        return self.world

    def get_rank(self, group=None):
        return 0

    def barrier(self, *a, **k):
        return None

    def all_reduce(self, t, *a, **k):
        return None

    def all_gather_object(self, out, obj, group=None):
        for i in range(len(out)):
            out[i] = obj

    def all_gather(self, outs, inp, group=None):
        for o in outs:
            o.copy_(inp)

    class _Done:
        def wait(self):
            return None

    def all_gather_into_tensor(self, out, inp, group=None, async_op=False):
        out.view(self.world, -1).copy_(inp.reshape(1, -1).expand(self.world, -1))
        return self._Done() if async_op else None

    isend, irecv = "isend", "irecv"

    def P2POp(self, op, tensor, peer, group=None):
        return (op, tensor, peer)

    def batch_isend_irecv(self, ops):
        sent = [t for op, t, _ in ops if op == "isend"]
        for op, t, _ in ops:
            if op == "irecv":
                t.copy_(sent[0])
        return [self._Done() for _ in ops]

    def destroy_process_group(self):
        self.up = False


def _run_bench(monkeypatch, capsys, argv, world=1):
    import bench
    from compactfusion_b200 import _native as nv
    from compactfusion_b200 import engine as eng_mod
    FakeCuda.reset()
    lib = IpcLib({})
    monkeypatch.setattr(nv, "lib", lambda: lib)
    monkeypatch.setattr(nv, "stream_ptr", lambda: FakeCuda.current().cuda_stream)
    monkeypatch.setattr(nv, "workspace", lambda nbytes, device: torch.empty(max(int(nbytes), 16), dtype=torch.uint8))
    monkeypatch.setattr(nv, "workspace_bytes", lambda *a, **k: 4096)
    monkeypatch.setattr(nv, "require_cuda_half", lambda t, name: None)
    # a made-up IPC pointer cannot be wrapped: the parity leg reads zeros instead of a receive slot
    monkeypatch.setattr(eng_mod, "_device_bytes", lambda ptr, nbytes, device: torch.zeros(nbytes, dtype=torch.uint8))
    for name, val in dict(current_stream=lambda *a, **k: FakeCuda.current(), Stream=FakeStream, Event=TimedEvent,
                          stream=FakeCuda.stream_ctx, is_available=lambda: True, set_device=lambda d: None,
                          synchronize=lambda *a, **k: None, CUDAGraph=FakeGraph, graph=fake_graph_ctx,
                          current_device=lambda: 0).items():
        monkeypatch.setattr(torch.cuda, name, val)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    # inside bench.py, "cuda" devices are the CPU
    real_device = torch.device
    proxy = types.SimpleNamespace(**{k: getattr(torch, k) for k in dir(torch) if not k.startswith("__")})
    proxy.device = lambda *a, **k: real_device("cpu")
    monkeypatch.setattr(bench, "torch", proxy)
    if world > 1:
        fd = FakeDist(world)
        monkeypatch.setattr(bench, "dist", fd)
        monkeypatch.setattr(eng_mod, "dist", fd)
        monkeypatch.setenv("WORLD_SIZE", str(world))
        monkeypatch.setenv("RANK", "0")
        monkeypatch.setenv("LOCAL_RANK", "0")
    else:
        for v in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
            monkeypatch.delenv(v, raising=False)
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    # restore the module-level workload constants afterwards
    saved = {k: getattr(bench, k) for k in ("LAYERS", "SEQ", "CH", "METRIC", "WORKLOAD", "MODE")}
    try:
        bench.main()
    finally:
        for k, v in saved.items():
            setattr(bench, k, v)
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "bench.py must print exactly ONE JSON line"
    return json.loads(lines[0])


BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"}


@pytest.mark.parametrize("argv", [
    ["--layers", "2", "--steps", "4", "--no-cpu-baseline", "--no-gpu-reference", "--no-parity"],
    ["--layers", "6", "--steps", "3", "--no-cpu-baseline", "--no-gpu-reference", "--overlap", "--codec", "int2", "--no-parity"],
    ["--layers", "2", "--steps", "3", "--no-cpu-baseline", "--no-gpu-reference", "--no-graph", "--no-e2e", "--no-parity"],
    ["--layers", "2", "--steps", "3", "--no-cpu-baseline", "--no-gpu-reference", "--codec", "raw"],
    ["--layers", "2", "--steps", "3", "--no-cpu-baseline", "--no-gpu-reference", "--no-e2e", "--workload", "pixart_patch_parallel"],
])
def test_bench_gpu_arm_single_gpu_dry_run(monkeypatch, capsys, argv):
    line = _run_bench(monkeypatch, capsys, argv)
    assert BASE_KEYS <= set(line), BASE_KEYS - set(line)
    assert line["n_gpus"] == 1 and line["unit"] == "GB/s" and line["higher_is_better"] is True
    assert line["steps"] == int(argv[argv.index("--steps") + 1]) and line["warmup"] >= 3
    cfg = line["config"]
    assert cfg["workload"] and "model" not in cfg and cfg["layers"] == int(argv[argv.index("--layers") + 1])
    raw = "raw" in argv
    if raw:
        assert line["impl"] == "uncompressed_baseline" and line["roofline"] is None
    else:
        r = line["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "kernels"} <= set(r)
        names = [k["kernel"] for k in r["kernels"]]
        assert names[0].startswith("k_delta_stats") and names[-1].startswith("k_apply_codes")
        assert ("k_int2_encode_tma" in names) == ("int2" in argv)
        assert line["gpu_launches"] > 0
    assert cfg["schedule"] == ("two chains (compress | reconstruct)" if "--overlap" in argv else "serial")
    assert (line["e2e"] is None) == ("--no-e2e" in argv)
    if line["e2e"]:
        assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert {"rel_l2", "max_abs", "psnr_db", "finite"} <= set(line["fidelity"]) and line["fidelity"]["finite"] is True
    assert line["ranks_identical"]["ok"] is True and isinstance(line["parity_ok"], bool)
    if "--no-parity" in argv or raw:
        assert line["oracle_parity"] is None
    else:  # the stand-in library computes nothing: the oracle must notice (and the line must survive it)
        assert line["oracle_parity"]["ok"] is False and line["parity_ok"] is False


def test_bench_config1_int4_roundtrip_dry_run(monkeypatch, capsys):
    """BASELINE configs[0] (one 4096 x 3072 tensor, INT4 + error feedback, 28-step series) through the plugin API:
    the line keeps the contract's keys, names the workload and counts the launches of the per-call path."""
    line = _run_bench(monkeypatch, capsys, ["--workload", "config1_int4_roundtrip", "--steps", "2", "--no-cpu-baseline"])
    assert BASE_KEYS <= set(line) and line["n_gpus"] == 1 and line["unit"] == "GB/s"
    cfg = line["config"]
    assert cfg["workload"] == "config1_int4_roundtrip" and cfg["codec"] == "int4" and cfg["series_steps"] == 28
    assert cfg["seq"] == 4096 and cfg["channels"] == 3072 and "model" not in cfg
    assert line["gpu_launches"] > 0 and line["roofline"]["bound"] == "hbm" and line["roofline"]["algorithmic_bytes_per_launch"] > 0
    assert line["e2e"] is None and line["cpu_baseline"] is None


@pytest.mark.parametrize("argv", [
    ["--gpus", "2", "--layers", "3", "--steps", "3", "--no-cpu-baseline", "--no-gpu-reference"],
    ["--gpus", "2", "--layers", "6", "--steps", "3", "--no-cpu-baseline", "--no-gpu-reference", "--no-e2e", "--overlap", "--no-parity"],
    ["--gpus", "2", "--layers", "3", "--steps", "3", "--no-cpu-baseline", "--no-gpu-reference", "--no-e2e", "--transport", "nccl", "--no-parity"],
    ["--gpus", "2", "--layers", "3", "--steps", "3", "--no-e2e", "--codec", "raw", "--raw-exchange", "ring"],
    ["--gpus", "2", "--layers", "3", "--steps", "3", "--no-e2e", "--codec", "raw", "--raw-exchange", "async"],
    ["--gpus", "2", "--layers", "3", "--steps", "3", "--no-cpu-baseline", "--no-gpu-reference", "--no-e2e", "--workload", "cogvideox5b_ring",
     "--codec", "int2", "--no-parity"],
])
def test_bench_gpu_arm_two_rank_dry_run(monkeypatch, capsys, argv):
    """Rank 0 of a pretended 2-rank job: the transport probe, CUDA-IPC region setup, fused put, flag-waiting
    reconstruct and the per-kernel legs of the one-sided path."""
    if "cogvideox5b_ring" in argv:
        import bench
        monkeypatch.setitem(bench.WORKLOADS, "cogvideox5b_ring", dict(bench.WORKLOADS["cogvideox5b_ring"], rows=2 * 1024))
    line = _run_bench(monkeypatch, capsys, argv, world=2)
    assert BASE_KEYS <= set(line) and line["n_gpus"] == 2
    cfg = line["config"]
    if "raw" in argv:
        assert line["impl"] == "uncompressed_baseline" and line["roofline"] is None and line["gpu_launches"] == 0
        assert cfg["launch_mode"] == "eager (uncompressed %s)" % argv[-1]
        return
    nccl = "nccl" in argv
    assert cfg["transport"].startswith("nccl" if nccl else "p2p"), cfg["transport"]
    assert cfg["launch_mode"] == ("eager" if nccl else "cuda_graph")
    assert line["cpu_baseline"] is None and line["p2p_wait_timeouts"] is False
    names = [k["kernel"] for k in line["roofline"]["kernels"]]
    if not nccl:
        assert all(k.get("fused_put") for k in line["roofline"]["kernels"][:2])
        assert "k_p2p_put" not in names
    if "cogvideox5b_ring" in argv:
        assert cfg["exchange"] == "ring"


@pytest.mark.parametrize("gpus", [1, 2])
def test_reference_arm_runs_on_the_host(gpus):
    """`bench.py --impl reference`: the reference's own eager-torch path when baseline/_ref is staged (kind
    "reference"), the oracle port otherwise; ms_per_step is the measured sample step, never an extrapolation."""
    import os
    import subprocess
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    t0 = time.time()
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", str(gpus),
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600,
                       env={k: v for k, v in os.environ.items() if k not in ("WORLD_SIZE", "RANK", "LOCAL_RANK")})
    wall = time.time() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == gpus and line["unit"] == "GB/s"
    staged = os.path.isdir(os.path.join(root, "baseline", "_ref", "xfuser", "compact"))
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"] and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["ms_per_step"] * line["steps"] / 1e3 < wall, "the claimed steps must fit inside the run"
    assert line["config"]["workload"] == "flux1024_patch_parallel"
