"""Development aid: where the end-to-end step of bench.py spends its time (host clock around every call)."""
import os, sys, time, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
logging.disable(logging.CRITICAL)
from niwqg_b200 import CoupledModel, InitialConditions as ic
from cases import lamb_params
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
kw, U0, k0 = lamb_params(nx, True, 1, 1)
kw["tmax"] = 1e30; kw["twrite"] = 10 ** 9
m = CoupledModel.Model(**kw)
q_pin = torch.empty((nx, nx), dtype=torch.float64).pin_memory()
phi_pin = torch.empty((nx, nx), dtype=torch.complex128).pin_memory()
qo = torch.empty((nx, nx), dtype=torch.float64).pin_memory()
po = torch.empty((nx, nx), dtype=torch.complex128).pin_memory()
q_pin.numpy()[...] = ic.LambDipole(m, U=U0, R=2 * np.pi / k0)
phi_pin.numpy()[...] = (1 + 1j) * (2 * U0) / np.sqrt(2)
h = m._h
def T(f, sync=False):
    t0 = time.perf_counter(); f()
    if sync: h.sync()
    return (time.perf_counter() - t0) * 1e3
for it in range(4):
    a = T(lambda: m.set_q(q_pin.numpy()))
    b = T(lambda: m.set_phi(phi_pin.numpy()))
    c0 = T(lambda: m._step_etdrk4())
    c1 = T(lambda: m._calc_derived_fields())
    d = T(lambda: h.field_into("Q", qo.numpy(), 0, wait=False))
    e = T(lambda: h.field_into("PHI", po.numpy(), 0, wait=False))
    f = T(lambda: h.sync())
    print("iter %d: set_q %.1f  set_phi %.1f  step(issue) %.1f  diagnostics %.1f  get Q %.1f  get PHI %.1f  final sync %.1f  total %.1f ms"
          % (it, a, b, c0, c1, d, e, f, a + b + c0 + c1 + d + e + f), flush=True)
# the bench loop: no synchronisation between iterations
h.sync(); t0 = time.perf_counter()
for it in range(5):
    m.set_q(q_pin.numpy()); m.set_phi(phi_pin.numpy()); m._step_forward()
    h.field_into("Q", qo.numpy(), 0, wait=False); h.field_into("PHI", po.numpy(), 0, wait=False)
h.sync()
print("pipelined loop: %.1f ms per iteration" % ((time.perf_counter() - t0) * 1e3 / 5))
# synchronous pieces
print("sync'd: set_q %.1f set_phi %.1f step %.1f diag %.1f getQ %.1f getPHI %.1f" % (
    T(lambda: m.set_q(q_pin.numpy()), True), T(lambda: m.set_phi(phi_pin.numpy()), True), T(lambda: m._step_etdrk4(), True),
    T(lambda: m._calc_derived_fields(), True), T(lambda: h.field_into("Q", qo.numpy(), 0), True), T(lambda: h.field_into("PHI", po.numpy(), 0), True)))
