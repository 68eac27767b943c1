"""Run the engines' Python control flow on the CPU against a RECORDING stand-in for CUDA streams / events and
for the native library (TEST ONLY: nothing is computed), turn what was recorded into the per-step dependency
graph -- which library call went to which stream, behind which events -- and feed that graph to the protocol
model checker (test_exchange_protocol_model.simulate).

This ties the model check to the schedule engine.py actually builds (serial step, ring order, two-chain
step) instead of a restatement of it, and exercises the host code of the multi-rank paths (pointer tables,
per-origin argument caches, launch counting) without a GPU.
"""
import contextlib

import pytest
import torch

from test_exchange_protocol_model import Violation, simulate

N_LOCAL, CH = 64, 256


class FakeEvent:
    def __init__(self, *a, **k):
        self.pos = None  # (stream, index of the last op recorded before the event)

    def record(self, stream=None):
        stream = stream or FakeCuda.current()
        self.pos = (stream, len(stream.ops) - 1)


class FakeStream:
    def __init__(self, *a, name=None, **k):
        self.name = name or f"side{len(FakeCuda.streams)}"
        self.ops = []      # dicts {kind: 'kernel', name, layer, origins, waits: [(stream, idx)]}
        self.pending_waits = []
        self.cuda_stream = 0x5000 + len(FakeCuda.streams)
        FakeCuda.streams.append(self)

    def wait_event(self, ev):
        if ev.pos is not None and ev.pos[1] >= 0:
            self.pending_waits.append(ev.pos)

    def wait_stream(self, other):
        if other.ops:
            self.pending_waits.append((other, len(other.ops) - 1))

    def launch(self, name, layer, origins):
        self.ops.append(dict(name=name, layer=layer, origins=origins, waits=self.pending_waits))
        self.pending_waits = []


class FakeCuda:
    streams, stack = [], []

    @classmethod
    def reset(cls):
        cls.streams, cls.stack = [], []
        cls.stack.append(FakeStream(name="main"))

    @classmethod
    def current(cls):
        return cls.stack[-1]

    @classmethod
    @contextlib.contextmanager
    def stream_ctx(cls, s):
        cls.stack.append(s)
        try:
            yield
        finally:
            cls.stack.pop()


class FakeLib:
    """Stands in for libcompactb200.so: every entry point returns 0 and records itself on the current stream."""

    def __init__(self, ctx):
        self.ctx = ctx

    def __getattr__(self, name):
        def call(*args):
            if name in ("cf_sign_compress_put", "cf_sign_compress_passes", "cf_p2p_put", "cf_sign_decompress_batched_wait"):
                FakeCuda.current().launch(name, self.ctx.get("layer"), self.ctx.get("origins"))
            return 0
        return call


@pytest.fixture
def fake_cuda(monkeypatch):
    from compactfusion_b200 import _native as nv
    ctx = {}
    FakeCuda.reset()
    lib = FakeLib(ctx)
    monkeypatch.setattr(nv, "lib", lambda: lib)
    monkeypatch.setattr(nv, "stream_ptr", lambda: FakeCuda.current().cuda_stream)
    monkeypatch.setattr(nv, "workspace", lambda nbytes, device: torch.empty(max(int(nbytes), 16), dtype=torch.uint8))
    monkeypatch.setattr(nv, "workspace_bytes", lambda *a, **k: 4096)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: FakeCuda.current())
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "stream", FakeCuda.stream_ctx)
    return ctx


def _engine(cls, layers, world, rank, ctx, ctype):
    """An engine of a `world`-rank job on fake device pointers (no process group, no CUDA IPC)."""
    eng = cls(layers, N_LOCAL, CH, device=torch.device("cpu"), transport="nccl")
    eng.world, eng.rank = world, rank
    eng.global_k = [torch.zeros((world * N_LOCAL, CH), dtype=torch.half) for _ in range(layers)]
    eng.global_v = [torch.zeros((world * N_LOCAL, CH), dtype=torch.half) for _ in range(layers)]
    if world > 1:
        eng.transport = "p2p"
        slot_bytes = 2 * eng._numel(ctype) * 2
        eng._p2p[ctype] = {"base": 0x10000000 * (rank + 1), "peers": [0x10000000 * (q + 1) for q in range(world)],
                           "flags_bytes": 256, "slot_bytes": slot_bytes,
                           "count": torch.zeros(layers, dtype=torch.int32), "ticket": torch.zeros(1, dtype=torch.int32),
                           "error": torch.zeros(1, dtype=torch.int32)}
    # tag every library call with the layer / origins of the engine method that issued it
    for meth in ("compress_put", "compress", "gather", "decompress"):
        orig = getattr(eng, meth)

        def wrapped(*a, _orig=orig, _meth=meth, **k):
            if _meth == "gather":
                ctx["layer"] = a[1] if len(a) > 1 else k.get("layer", 0)
            else:
                ctx["layer"] = a[0]
            ctx["origins"] = (k.get("origins") or (a[2] if len(a) > 2 else None)) if _meth == "decompress" else None
            return _orig(*a, **k)
        setattr(eng, meth, wrapped)
    return eng


def _record_step(cls, layers, world, rank, ctx, ctype, overlap):
    FakeCuda.reset()
    eng = _engine(cls, layers, world, rank, ctx, ctype)
    x = [torch.zeros((N_LOCAL, CH), dtype=torch.half) for _ in range(layers)]
    eng.step(x, x, ctype, overlap)
    return eng


def _template(world):
    """Recorded streams -> model template.  A layer's put = the LAST compress-side launch of that layer
    (the kernel that publishes the flags); apply = every flag-waiting reconstruct launch."""
    flat, index = [], {}
    # order entries so that dependencies point backwards: interleave by recording order is not kept per stream,
    # so resolve in passes
    pending = [(s, i) for s in FakeCuda.streams for i in range(len(s.ops))]
    last_put_of = {}
    for s in FakeCuda.streams:
        for i, op in enumerate(s.ops):
            if op["name"] != "cf_sign_decompress_batched_wait":
                last_put_of[op["layer"]] = (s, i)
    while pending:
        progressed = False
        for key in list(pending):
            s, i = key
            op = s.ops[i]
            deps = list(op["waits"]) + ([(s, i - 1)] if i > 0 else [])
            if any(d not in index for d in deps):
                continue
            is_apply = op["name"] == "cf_sign_decompress_batched_wait"
            kind = "apply" if is_apply else ("put" if last_put_of[op["layer"]] == key else "local")
            origins = tuple(op["origins"]) if (is_apply and op["origins"]) else (tuple(range(world)) if is_apply else None)
            index[key] = len(flat)
            flat.append(dict(kind=kind, l=op["layer"], stream=s.name, origins=origins,
                             deps=[index[d] for d in deps if d[0] is not s]))
            pending.remove(key)
            progressed = True
        assert progressed, "cyclic dependencies in the recorded schedule"
    return flat


def _simulate_recorded(templates, world, layers, seeds=40):
    # 'local' entries (stats / finalize kernels before the publishing one) only carry ordering: model them as
    # applies without origins (no slot access)
    for tpl in templates:
        for e in tpl:
            if e["kind"] == "local":
                e["kind"], e["origins"] = "apply", ()
    for seed in range(seeds):
        simulate(world, layers, steps=3, mode=None, lag=0, seed=seed, templates=templates)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("schedule", ["serial", "overlap", "ring", "serial_putkernel"])
def test_recorded_engine_schedules_pass_the_model_check(fake_cuda, monkeypatch, world, schedule):
    from compactfusion_b200.engine import PatchGatherEngine, RingExchangeEngine
    from compactfusion_b200.utils import COMPACT_COMPRESS_TYPE as T
    layers = 6
    if schedule == "serial_putkernel":
        monkeypatch.setenv("CF_FUSED_PUT", "0")
    cls = RingExchangeEngine if schedule == "ring" else PatchGatherEngine
    templates = []
    for rank in range(world):
        eng = _record_step(cls, layers, world, rank, fake_cuda, T.BINARY, overlap=schedule == "overlap")
        tpl = _template(world)
        templates.append(tpl)
        applies = [e for e in tpl if e["kind"] == "apply"]
        puts = [e for e in tpl if e["kind"] == "put"]
        assert len(puts) == layers and sorted(e["l"] for e in puts) == list(range(layers))
        if schedule == "ring":
            assert len(applies) == layers * world
            assert [e["origins"] for e in applies[:world]] == [((rank - h) % world,) for h in range(world)]
        else:
            assert len(applies) == layers and all(e["origins"] == tuple(range(world)) for e in applies)
        streams = {e["stream"] for e in tpl}
        assert streams == ({"main", "side1"} if schedule == "overlap" else {"main"}), streams
        if schedule == "overlap":
            assert all(e["stream"] == "side1" for e in applies) and all(e["stream"] == "main" for e in puts)
            assert eng.kernel_launches == layers * 4  # stats + finalize (fused put) + flag publication + reconstruct
        elif schedule == "serial_putkernel":
            assert eng.kernel_launches == layers * 4  # stats, finalize, put kernel, reconstruct
        else:
            assert eng.kernel_launches == layers * (3 + (world if schedule == "ring" else 1))  # + k_publish_flags
    _simulate_recorded(templates, world, layers)


def test_recorded_overlap_schedule_with_a_broken_lag_is_caught(fake_cuda, monkeypatch):
    """Same harness, OVERLAP_LAG raised past the layer count: the recorded graph loses the put -> reconstruct
    back-pressure and the model check reports the slot race."""
    from compactfusion_b200.engine import PatchGatherEngine
    from compactfusion_b200.utils import COMPACT_COMPRESS_TYPE as T
    layers, world = 5, 2
    monkeypatch.setattr(PatchGatherEngine, "OVERLAP_LAG", 99)
    monkeypatch.setattr(PatchGatherEngine, "can_overlap", lambda self, ctype: True)
    templates = []
    for rank in range(world):
        _record_step(PatchGatherEngine, layers, world, rank, fake_cuda, T.BINARY, overlap=True)
        templates.append(_template(world))
    with pytest.raises(Violation):
        _simulate_recorded(templates, world, layers, seeds=300)


def test_single_gpu_overlap_uses_per_layer_payload_buffers(fake_cuda):
    """World 1 has no receive slots: recv aliases send, so the two-chain step must give every layer its own
    payload buffer (compress of layer l+1 runs beside the reconstruct of layer l)."""
    from compactfusion_b200.engine import PatchGatherEngine
    from compactfusion_b200.utils import COMPACT_COMPRESS_TYPE as T
    layers = 6
    eng = _record_step(PatchGatherEngine, layers, 1, 0, fake_cuda, T.INT2, overlap=True)
    assert eng._per_layer_send
    bufs = {eng._buffers(T.INT2, l)[0].data_ptr() for l in range(layers)}
    assert len(bufs) == layers
    assert all(eng._buffers(T.INT2, l)[1].data_ptr() == eng._buffers(T.INT2, l)[0].data_ptr() for l in range(layers))
    tpl = _template(1)
    assert {e["stream"] for e in tpl} == {"main", "side1"}
    # serial engines keep ONE buffer for all layers
    eng2 = _record_step(PatchGatherEngine, layers, 1, 0, fake_cuda, T.INT2, overlap=False)
    assert len({eng2._buffers(T.INT2, l)[0].data_ptr() for l in range(layers)}) == 1
    # too few layers for the slot-reuse argument: the flag is ignored
    eng3 = _record_step(PatchGatherEngine, 4, 1, 0, fake_cuda, T.INT2, overlap=True)
    assert not eng3.can_overlap(T.INT2) and {e["stream"] for e in _template(1)} == {"main"}
