"""Quick device timing of the coupled step and of the bare FFT at several sizes (development aid)."""
import sys, os, time, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
logging.disable(logging.CRITICAL)
from niwqg_b200 import _native as nat

sizes = [int(a) for a in sys.argv[1:]] or [512, 2048, 8192]
for N in sizes:
    h = nat.Handle(model=nat.MODEL_COUPLED, nx=N, batch=1, device=0, L=2 * np.pi * 200e3, dt=1e4 * 128 / N, U=-0.1,
                   f=1e-4, N=0.01, m=2 * np.pi / 280, nu=20., nu4=5e11 * (128 / N) ** 4, nuw=50., use_filter=1)
    rng = np.random.RandomState(0)
    q = 1e-5 * rng.randn(N, N)
    h.set_q(q)
    h.set_phi((np.ones((N, N)) + 1j) * 0.14)
    h.time_steps(2)
    l0 = h.launch_count()
    n = 3 if N >= 4096 else 10
    ms = h.time_steps(n) / n
    l1 = h.launch_count()
    pts = N * N
    print("N=%5d coupled step %.3f ms  %.3e pt-steps/s  launches/step %d  eff(3392B) %.1f%% of 6544 GB/s"
          % (N, ms, pts / (ms * 1e-3), (l1 - l0) // n, 100 * 3392 * pts / (ms * 1e-3) / 6544e9), flush=True)
    h.close()
