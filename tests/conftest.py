import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "multigpu(n): needs n GPUs in one box (deselected on smaller boxes)")


def pytest_collection_modifyitems(config, items):
    """Tests marked `multigpu(n)` need n GPUs in one box: on a smaller box they are DESELECTED (their world-size-n
    logic is covered there by the virtual-rank tests of test_gpu_engine.py; tools/gpu_multi.sh runs them on
    2 / 4 / 8 GPUs), so a 1-GPU `pytest -m gpu` run reports no skips."""
    have = torch.cuda.device_count() if torch.cuda.is_available() else 0
    keep, drop = [], []
    for it in items:
        m = it.get_closest_marker("multigpu")
        (drop if (m is not None and have < int(m.args[0])) else keep).append(it)
    if drop:
        config.hook.pytest_deselected(items=drop)
        items[:] = keep


def pytest_sessionstart(session):
    """The shared library is a build artefact (git-ignored): if this checkout does not have it yet, compile it
    once before any test maps it (nvcc, sm_100a; __graft_entry__.build() does the same).  This is not a
    fallback: without nvcc the build raises and every test that needs the library fails loudly."""
    from compactfusion_b200 import build as cf_build
    if not os.path.exists(cf_build.OUT):
        cf_build.build()


def free_port() -> str:
    """A TCP port that is free right now (for the rendezvous of multi-process tests)."""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sock:
        sock.bind(("127.0.0.1", 0))
        return str(sock.getsockname()[1])


def h16(a: np.ndarray) -> torch.Tensor:
    """uint16 bit pattern array -> fp16 CPU tensor (inverse of oracle/make_goldens.bits)."""
    return torch.from_numpy(a.view(np.int16).copy()).view(torch.half)


def bits16(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().contiguous().view(torch.int16).numpy().view(np.uint16)


def assert_bits_equal(a: torch.Tensor, b: torch.Tensor, what=""):
    """Bit-exact fp16 equality (NaN payloads included)."""
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    x, y = bits16(a), bits16(b)
    bad = int((x != y).sum())
    assert bad == 0, f"{what}: {bad} of {x.size} fp16 values differ bitwise"


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.float().cpu(), b.float().cpu()
    return float(torch.norm(a - b) / torch.norm(b).clamp_min(1e-30))


@pytest.fixture(scope="session")
def golden_codecs():
    return np.load(os.path.join(GOLDEN, "codecs.npz"))


@pytest.fixture(scope="session")
def golden_slowpath():
    return np.load(os.path.join(GOLDEN, "slowpath.npz"))


@pytest.fixture(scope="session")
def golden_lowrank_q_wire():
    """LOW_RANK_Q payloads the reference assembled from fixed factors, and its reconstructions of them
    (oracle/make_goldens.py lowrank_q_wire_goldens)."""
    return np.load(os.path.join(GOLDEN, "lowrank_q_wire.npz"))


LRQ_WIRE_CASES = [("n64_c256_r4", 64, 256, 4), ("n130_c264_r20", 130, 264, 20), ("n96_c512_r32", 96, 512, 32)]


@pytest.fixture(scope="session")
def golden_state():
    return np.load(os.path.join(GOLDEN, "state_machine.npz"))


CODEC_CASES = ["rand_64x256", "rand_48x1152", "rand_130x64", "flux_k_96x512"]
