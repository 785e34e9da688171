"""Per-kernel-kind CUDA-event breakdown of the coupled step at several sizes (development aid)."""
import sys, os, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
logging.disable(logging.CRITICAL)
from niwqg_b200 import _native as nat

sizes = [a for a in sys.argv[1:]] or ["2048", "8192"]
for spec in sizes:
    N, B = (int(x) for x in (spec.split("x") + ["1"])[:2])
    h = nat.Handle(model=nat.MODEL_COUPLED, nx=N, batch=B, device=0, L=2 * np.pi * 200e3, dt=1e4 * 128 / N, U=-0.1,
                   f=1e-4, N=0.01, m=2 * np.pi / 280, nu=20., nu4=5e11 * (128 / N) ** 4, nuw=50., use_filter=1)
    rng = np.random.RandomState(0)
    h.set_q(1e-5 * rng.randn(B, N, N) if B > 1 else 1e-5 * rng.randn(N, N))
    h.set_phi((np.ones((B, N, N) if B > 1 else (N, N)) + 1j) * 0.14)
    h.time_steps(2)
    n = 3
    ms = h.time_steps(n) / n
    h.profile(True)
    h.step(n)
    prof = h.profile(False)
    pts = N * N * B
    print("N=%d B=%d step %.3f ms (%.1f%% of 3392B roofline)" % (N, B, ms, 100 * 3392 * pts / (ms * 1e-3) / 6544e9))
    for k, (t, c) in prof.items():
        if c:
            per = t / c
            gbs = 32.0 * pts / (per * 1e-3) / 1e9 if k.startswith("fft") else float("nan")
            print("   %-8s %8.3f ms/step  %4d launches/step  %.4f ms each  %7.0f GB/s(32B/pt)" % (k, t / n, c // n, per, gbs))
    h.close()
