"""Slab decomposition on real GPUs (needs >= 2 devices; skipped on a one-GPU box): the slab run must reproduce the
single-GPU run of the same model - q, phi, qh, Ke/Pw/Kw and diagnostics - to 1e-10 (observed ~1e-16), through
tools/slab_check.py launched with torch.distributed.run."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["256 5 coupled", "256 5 uncoupled", "256 5 ybj", "256 5 ql", "1024 3 coupled"])
def test_slab_matches_single_gpu(cfg):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "slab_check.py")] + cfg.split()
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.gpu
def test_handles_on_two_devices_of_one_process():
    """The > 48 KB dynamic shared-memory opt-in of the FFT kernels is per device: a handle created on a second device
    after one on the first must still launch (it used to be guarded by a process-wide flag)."""
    if _ngpu() < 2:
        pytest.skip("needs at least 2 GPUs")
    import numpy as np
    from niwqg_b200 import _native as nat
    rng = np.random.RandomState(5)
    for N in (2048,):
        x = rng.randn(N, N) + 1j * rng.randn(N, N)
        ref = np.fft.fft2(x)
        for dev in (0, 1):
            h = nat.Handle(model=nat.MODEL_UNCOUPLED, nx=N, batch=1, device=dev, L=5e5, dt=1e4, f=1e-4, N=0.01, m=0.025,
                           nu=20., nuw=50.)
            X = h.fft2(x, nat.FFT_C2C_FWD)
            assert np.linalg.norm(X - ref) / np.linalg.norm(ref) < 2e-15, dev
            h.close()
