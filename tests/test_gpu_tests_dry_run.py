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
    # the product's fused LSE merge refuses CPU tensors: the dry run merges with the eager restatement
    from compactfusion_b200 import attention

    def eager_merge(out, lse, block_out, block_lse):
        if out is None:
            return block_out.to(torch.float32), block_lse.contiguous().to(torch.float32)
        o, l = attention.update_out_and_lse(out, lse.transpose(1, 2).unsqueeze(-1), block_out, block_lse)
        return o, l.squeeze(-1).transpose(1, 2).contiguous()
    monkeypatch.setattr(attention, "merge_out_and_lse", eager_merge)


@pytest.mark.parametrize("codec", ["binary", "int2"])
def test_dry_run_ring_engine_world1(fake_gpu, codec):
    zz.test_ring_engine_world1_equals_patch_engine_and_plain_attention(codec)


@pytest.mark.parametrize("codec", ["binary", "int2"])
def test_dry_run_overlapped_step_world1(fake_gpu, codec):
    zz.test_overlapped_step_equals_serial_world1(codec)  # (a direct call ignores the CF_EXPERIMENTAL skip mark)


@pytest.mark.parametrize("worker", ["WORKER", "OVERLAP_WORKER"])
def test_dry_run_two_gpu_workers_as_rank0(fake_gpu, monkeypatch, worker):
    """The 2-GPU worker scripts of test_gpu_zz_ring_engine, executed in-process as rank 0 of a pretended
    2-rank job (fake process group, fake CUDA IPC): transports, fused put, per-hop reconstruct, graph capture
    and the launch-count assertions."""
    import torch.distributed as dist
    from compactfusion_b200 import _native as nv
    from test_bench_dry_run import FakeDist, IpcLib
    lib = IpcLib({})
    monkeypatch.setattr(nv, "lib", lambda: lib)
    fd = FakeDist(2)
    for name in ("init_process_group", "is_initialized", "get_world_size", "get_rank", "barrier", "all_reduce",
                 "all_gather_object", "all_gather_into_tensor", "destroy_process_group"):
        monkeypatch.setattr(dist, name, getattr(fd, name))
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("CF_ROOT", zz.ROOT)
    src = getattr(zz, worker).replace('torch.device("cuda", rank)', 'torch.device("cpu")')
    src = src.replace('device_id=dev', 'device_id=None')
    out = {}
    exec(compile(src, worker, "exec"), {"__name__": "__worker__", "print": lambda *a: out.setdefault("printed", a)})
    assert out["printed"] == ("WORKER_OK", 0)
