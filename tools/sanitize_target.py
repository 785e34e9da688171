"""compute-sanitizer target: one forward + inverse transform and two coupled steps at N (cluster kernels for N >= 2048)."""
import sys, os, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
logging.disable(logging.CRITICAL)
from niwqg_b200 import _native as nat
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
h = nat.Handle(model=nat.MODEL_COUPLED, nx=N, batch=1, device=0, L=2 * np.pi * 200e3, dt=1e4 * 128 / N, U=-0.1,
               f=1e-4, N=0.01, m=2 * np.pi / 280, nu=20., nu4=5e11 * (128 / N) ** 4, nuw=50., use_filter=1)
rng = np.random.RandomState(0)
x = rng.randn(N, N) + 1j * rng.randn(N, N)
X = h.fft2(x, nat.FFT_C2C_FWD)
xb = h.fft2(X, nat.FFT_C2C_INV)
print("round trip", float(np.abs(xb - x).max()))
h.set_q(1e-5 * rng.randn(N, N)); h.set_phi((np.ones((N, N)) + 1j) * 0.14)
h.step(2); h.sync()
print("Ke", h.scalars()[0, 0])
h.close()
