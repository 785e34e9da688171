import sys, os, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
logging.disable(logging.CRITICAL)
from cases import CASES, lamb_params, load_golden, rel_l2
from oracle import niwqg_oracle as orc
from test_gpu_parity import build_cuda

# 1. tables at 128 vs oracle
kw, U0, k0 = lamb_params(128, True, 1000, 100)
from niwqg_b200 import CoupledModel
m = CoupledModel.Model(**kw)
o = orc.NIWQGOracle(model="coupled", **kw)
for n in ["expch", "expch_h", "Qh", "f0", "fab", "fc", "expchw", "expch_hw", "Qhw", "f0w", "fabw", "fcw", "filtr"]:
    a, b = getattr(m, n), getattr(o, n)
    d = np.abs(a - b)
    print("table %-8s max abs diff/max|ref| %.2e   max rel (|ref|>1e-300) %.2e  #rel>1e-12: %d" % (
        n, d.max() / np.abs(b).max(), (d / np.maximum(np.abs(b), 1e-300)).max(), int(((d / np.maximum(np.abs(b), 1e-300)) > 1e-12).sum())))
# 2. oracle with device tables vs oracle
q = orc.lamb_dipole(o, U=U0, R=2 * np.pi / k0); phi = (np.ones_like(q) + 1j) * (2 * U0) / np.sqrt(2)
o2 = orc.NIWQGOracle(model="coupled", **kw)
for n in ["expch", "expch_h", "Qh", "f0", "fab", "fc", "expchw", "expch_hw", "Qhw", "f0w", "fabw", "fcw", "filtr"]:
    setattr(o2, n, getattr(m, n))
for mdl in (o, o2, m):
    mdl.set_q(q); mdl.set_phi(phi)
for s in range(1, 101):
    o.step(); o2.step(); m._step_etdrk4()
    if s in (1, 2, 5, 10, 20, 50, 100):
        print("step %3d  oracle(dev tables) vs oracle: q %.2e phi %.2e | cuda vs oracle: q %.2e phi %.2e | cuda vs oracle(dev tables): q %.2e phi %.2e" % (
            s, rel_l2(o2.q, o.q), rel_l2(o2.phi, o.phi), rel_l2(m.q, o.q), rel_l2(m.phi, o.phi), rel_l2(m.q, o2.q), rel_l2(m.phi, o2.phi)), flush=True)
# 3. diagnostics series detail
for name in ["uncoupled_lamb64_filt", "coupled_lamb64_filt"]:
    g = load_golden(name)
    mm = build_cuda(name)
    mm.run()
    for k, ref in g.items():
        if k.startswith("diag_"):
            got = np.asarray(mm.diagnostics[k[5:]]['value'], float)
            d = np.abs(got - ref)
            print("%s %-9s max|ref| %.3e max abs diff %.2e (rel %.1e)  at idx %d" % (name, k[5:], np.abs(ref).max(), np.nanmax(d), np.nanmax(d) / max(np.abs(ref).max(), 1e-300), int(np.nanargmax(d))))
