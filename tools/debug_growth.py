import sys, os, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
logging.disable(logging.CRITICAL)
from cases import lamb_params, rel_l2
from oracle import niwqg_oracle as orc
from niwqg_b200 import CoupledModel, UnCoupledModel, YBJModel, QGModel

def run(model, N=128, filt=True, over=None, steps=(1, 10, 50, 100)):
    qg = model == "qg"
    kw, U0, k0 = lamb_params(N, filt, 100000, 100, qg=qg)
    kw.update(over or {})
    if qg:
        m = QGModel.Model(**kw); o = orc.QGOracle(**kw)
    else:
        m = {"coupled": CoupledModel, "uncoupled": UnCoupledModel, "ybj": YBJModel}[model].Model(**kw)
        o = orc.NIWQGOracle(model=model, **kw)
    q = orc.lamb_dipole(o, U=U0, R=2 * np.pi / k0); phi = (np.ones_like(q) + 1j) * (2 * U0) / np.sqrt(2)
    for mdl in (m, o):
        mdl.set_q(q)
        if not qg: mdl.set_phi(phi)
    out = []
    for s in range(1, max(steps) + 1):
        m._step_etdrk4(); o.step()
        if s in steps:
            out.append("%d: q %.1e%s" % (s, rel_l2(m.q, o.q), "" if qg else " phi %.1e" % rel_l2(m.phi, o.phi)))
    print("%-10s N=%d filt=%d %s :: %s" % (model, N, filt, over or "", " | ".join(out)), flush=True)
    return m, o

run("qg"); run("ybj"); run("uncoupled"); m, o = run("coupled")
dq = np.abs(m.qh - o.qh) / np.abs(o.qh).max()
idx = np.argsort(dq.ravel())[::-1][:8]
print("largest |dqh|/max|qh| modes (ky,kx):", [(int(i // 128), int(i % 128), float("%.1e" % dq.ravel()[i])) for i in idx])
dp = np.abs(m.phih - o.phih) / np.abs(o.phih).max()
idx = np.argsort(dp.ravel())[::-1][:8]
print("largest |dphih|/max|phih| modes:", [(int(i // 128), int(i % 128), float("%.1e" % dp.ravel()[i])) for i in idx])
run("coupled", over={"U": 0.0})
run("coupled", over={"nu4": 0.0, "nu": 0.0, "nuw": 0.0})
run("coupled", filt=False)
run("coupled", N=64)
run("coupled", N=256, steps=(1, 10, 50))
