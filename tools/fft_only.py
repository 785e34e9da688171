"""Run a few bare 2-D FFTs of size N on the device (ncu target)."""
import sys, os, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
logging.disable(logging.CRITICAL)
from niwqg_b200 import _native as nat
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
h = nat.Handle(model=nat.MODEL_YBJ, nx=N, batch=1, device=0, L=5e5, dt=1e4, f=1e-4, N=0.01, m=0.025, nu=20., nuw=50.)
x = np.random.RandomState(0).randn(N, N) + 0j
for _ in range(2):
    X = h.fft2(x, nat.FFT_C2C_FWD)
h.close()
