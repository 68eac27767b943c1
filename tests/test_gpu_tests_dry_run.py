"""Host-code dry run of GPU tests that have not met a GPU yet: their bodies are executed here on CPU tensors with
the recording stand-ins for CUDA and the native library (test_engine_schedule_fake_cuda / test_bench_dry_run).
Nothing is reconstructed, so value assertions that compare two engines hold trivially; what this catches is
what would otherwise only surface on the GPU box -- wrong keyword names, tuple arities, launch-count
bookkeeping, shape asserts in ring_forward, the CPU attention reference path."""
import pytest
import torch

import test_gpu_zz_ring_engine as zz
from test_bench_dry_run import FakeGraph, fake_graph_ctx
from test_engine_schedule_fake_cuda import FakeCuda, FakeEvent, FakeLib, FakeStream


@pytest.fixture
def fake_gpu(monkeypatch):
    from compactfusion_b200 import _native as nv
    FakeCuda.reset()
    lib = FakeLib({})
    monkeypatch.setattr(nv, "lib", lambda: lib)
    monkeypatch.setattr(nv, "stream_ptr", lambda: FakeCuda.current().cuda_stream)
    monkeypatch.setattr(nv, "workspace", lambda nbytes, device: torch.empty(max(int(nbytes), 16), dtype=torch.uint8))
    monkeypatch.setattr(nv, "workspace_bytes", lambda *a, **k: 4096)
    for name, val in dict(current_stream=lambda *a, **k: FakeCuda.current(), Stream=FakeStream, Event=FakeEvent,
                          stream=FakeCuda.stream_ctx, synchronize=lambda *a, **k: None, CUDAGraph=FakeGraph,
                          graph=fake_graph_ctx, current_device=lambda: 0).items():
        monkeypatch.setattr(torch.cuda, name, val)
    monkeypatch.setattr(zz, "_cuda", lambda: torch.device("cpu"))


@pytest.mark.parametrize("codec", ["binary", "int2"])
def test_dry_run_ring_engine_world1(fake_gpu, codec):
    zz.test_ring_engine_world1_equals_patch_engine_and_plain_attention(codec)


@pytest.mark.parametrize("codec", ["binary", "int2"])
def test_dry_run_overlapped_step_world1(fake_gpu, codec):
    zz.test_overlapped_step_equals_serial_world1(codec)  # (a direct call ignores the CF_EXPERIMENTAL skip mark)
