"""ncu target: a couple of coupled steps at 8192^2 (fused split path)."""
import sys, os, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
logging.disable(logging.CRITICAL)
from niwqg_b200 import _native as nat
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
h = nat.Handle(model=nat.MODEL_COUPLED, nx=N, batch=1, device=0, L=2 * np.pi * 200e3, dt=1e4 * 128 / N, U=-0.1,
               f=1e-4, N=0.01, m=2 * np.pi / 280, nu=20., nu4=5e11 * (128 / N) ** 4, nuw=50., use_filter=1)
rng = np.random.RandomState(0)
h.set_q(1e-5 * rng.randn(N, N))
h.set_phi((np.ones((N, N)) + 1j) * 0.14)
h.step(nsteps)
h.sync()
ms = h.time_steps(2) / 2
print("N=%d step %.3f ms" % (N, ms))
h.close()
