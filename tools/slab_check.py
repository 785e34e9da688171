"""torchrun target: slab-decomposed run on WORLD_SIZE GPUs against a single-GPU run of the same model on rank 0's GPU.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/slab_check.py [nx] [nsteps] [model]"""
import os, sys, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import torch.distributed as dist
logging.disable(logging.CRITICAL)
from niwqg_b200 import CoupledModel, UnCoupledModel, YBJModel, QLModel, slab, _native as nat
from niwqg_b200 import InitialConditions as ic
from cases import lamb_params, rel_l2

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mname = sys.argv[3] if len(sys.argv) > 3 else "coupled"
cls = {"coupled": CoupledModel, "uncoupled": UnCoupledModel, "ybj": YBJModel, "ql": QLModel}[mname].Model
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
kw, U0, k0 = lamb_params(nx, True, 1, nsteps)
kw["twrite"] = 10 ** 9
m = slab.make_model(cls, dist=dist, **kw)
ref = cls(device=local, **kw)                 # the same model on one GPU (every rank runs its own copy)
q0 = ic.LambDipole(ref, U=U0, R=2 * np.pi / k0)
rng = np.random.RandomState(3)
q0 = q0 + 0.05 * np.abs(q0).max() * rng.randn(nx, nx)
phi0 = (np.ones_like(q0) + 1j) * (2 * U0) / np.sqrt(2) * (1 + 0.1 * rng.randn(nx, nx))
for mdl in (m, ref):
    mdl.set_q(q0); mdl.set_phi(phi0)
lo, hi = slab.rows_of(rank, world, nx)
# bare transform first
X = m._h.fft2(np.ascontiguousarray(phi0[lo:hi]), nat.FFT_C2C_FWD)
Xr = np.fft.fft2(phi0)[:, nat.slab_kx(nx, world, rank)]
e_fft = rel_l2(X, Xr)
xb = m._h.fft2(X, nat.FFT_C2C_INV)
e_ifft = rel_l2(xb, phi0[lo:hi])
for _ in range(nsteps):
    m._step_forward(); ref._step_forward()
eq, ep = rel_l2(m.q, ref.q[lo:hi]), rel_l2(m.phi, ref.phi[lo:hi])
# spectral slabs: a rank that only holds high wavenumbers has a tiny share of the spectrum's norm, so its rounding
# noise is measured against the whole spectrum (per rank), not against its own slab
refqh = ref.qh
eqh = float(np.linalg.norm(m.qh - refqh[:, nat.slab_kx(nx, world, rank)]) / (np.linalg.norm(refqh) / np.sqrt(world)))
sc = [abs(getattr(m, k) - getattr(ref, k)) / abs(getattr(ref, k)) for k in ("Ke", "Pw", "Kw")]
dg = max(abs(np.ravel(m.diagnostics[k]["value"])[-1] - np.ravel(ref.diagnostics[k]["value"])[-1]) /
         (abs(np.ravel(ref.diagnostics[k]["value"])[-1]) + 1e-300) for k in ("ke_qg", "ke_niw", "pe_niw", "ens"))
qg = slab.gather_rows(m, m.q, dist)
eg = rel_l2(qg, ref.q)
print("rank %d/%d %s nx=%d steps=%d: fft %.1e ifft %.1e | q %.2e phi %.2e qh %.2e gathered-q %.2e | Ke,Pw,Kw %.1e %.1e %.1e | diags %.1e"
      % (rank, world, mname, nx, nsteps, e_fft, e_ifft, eq, ep, eqh, eg, sc[0], sc[1], sc[2], dg), flush=True)
ok = max(e_fft, e_ifft) < 1e-13 and max(eq, ep, eqh, eg) < 1e-10 and max(sc) < 1e-10 and dg < 1e-9
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
