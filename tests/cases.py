"""Shared case builders for the parity tests: same parameters/ICs as tests/golden/make_golden.py."""
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    # name: (model, nx, use_filter, tdiags, nsteps, ic)
    "coupled_lamb64_filt": ("coupled", 64, True, 2, 10, "lamb"),
    "coupled_lamb64_nofilt": ("coupled", 64, False, 2, 10, "lamb"),
    "uncoupled_lamb64_filt": ("uncoupled", 64, True, 3, 10, "lamb"),
    "uncoupled_lamb64_nofilt": ("uncoupled", 64, False, 3, 10, "lamb"),
    "ql_lamb64_filt": ("ql", 64, True, 2, 10, "lamb"),
    "ql_lamb64_nofilt": ("ql", 64, False, 2, 10, "lamb"),
    "ybj_lamb64_filt": ("ybj", 64, True, 2, 10, "lamb"),
    "ybj_lamb64_nofilt": ("ybj", 64, False, 2, 10, "lamb"),
    "coupled_rand64_filt": ("coupled", 64, True, 5, 10, "mcw"),
    "ybj_rand64_filt": ("ybj", 64, True, 5, 10, "mcw"),
    "coupled_lamb128_nofilt_100": ("coupled", 128, False, 1, 100, "lamb"),
    "coupled_lamb128_filt_100": ("coupled", 128, True, 10, 100, "lamb"),
    "qg_lamb64_filt": ("qg", 64, True, 2, 20, "lamb"),
    "qg_lamb128_nofilt_100": ("qg", 128, False, 1, 100, "lamb"),
    "qg_scalar64_nofilt": ("qgc", 64, False, 1, 20, "lamb"),
    # parameter branches no other case touches (constructor keywords in EXTRA below)
    "coupled_lamb64_diss": ("coupled", 64, True, 1, 10, "lamb"),       # nu4w, mu, muw != 0 (Kernel.py:688-689, :646-652)
    "coupled_lamb64_dealias": ("coupled", 64, False, 2, 10, "lamb"),   # 2/3-rule mask (Kernel.py:277-281)
    "uncoupled_lamb64_diss": ("uncoupled", 64, True, 2, 10, "lamb"),   # nu4w != 0 without the wave PV
    "ql_lamb64_diss": ("ql", 64, True, 2, 10, "lamb"),                 # nu4w != 0 on the physical-space budget path
    "qg_lamb64_beta": ("qg", 64, True, 2, 20, "lamb"),                 # beta != 0 (QGModel.py:428)
}

EXTRA = {
    "coupled_lamb64_diss": dict(nu4w=2.e11, mu=2.e-8, muw=3.e-8),
    "coupled_lamb64_dealias": dict(dealias=True),
    "uncoupled_lamb64_diss": dict(nu4w=2.e11, muw=3.e-8),
    "ql_lamb64_diss": dict(nu4w=2.e11, mu=2.e-8, muw=3.e-8),
    "qg_lamb64_beta": dict(beta=2.e-11, mu=1.e-8),
}


def lamb_params(nx, use_filter, tdiags, nsteps, qg=False):
    L = 2 * np.pi * 200e3
    k0 = 10 * (2 * np.pi / L)
    U0 = 1.e-1
    Te = (U0 * k0) ** -1
    if qg:
        dt = .05 * Te * 128 / nx
        kw = dict(L=L, nx=nx, dt=dt, tmax=nsteps * dt - 0.5 * dt, twrite=7, nu4=7.5e8 * (128 / nx) ** 4,
                  use_filter=use_filter, U=-U0, tdiags=tdiags, beta=0.)
    else:
        dt = .025 * Te * 128 / nx
        kw = dict(L=L, nx=nx, dt=dt, tmax=nsteps * dt - 0.5 * dt, twrite=7, m=2 * np.pi / 280, N=0.01,
                  f=1.e-4, nu4=5e11 * (128 / nx) ** 4, nu4w=0., nu=20, nuw=50., mu=0., muw=0.,
                  use_filter=use_filter, U=-U0, tdiags=tdiags)
    return kw, U0, k0


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel_l2(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))
