"""Full-size checks of the CUDA path (BASELINE.json sizes): parity against the oracle at the largest sizes the numpy
oracle finishes in seconds, and size-independent properties at 8192^2 where no CPU reference can run (the
reference cannot even initialise above ~2048^2)."""
import logging

import numpy as np
import pytest

from cases import lamb_params, rel_l2

pytestmark = pytest.mark.gpu
logging.disable(logging.CRITICAL)


def _pair(model, nx, nsteps, use_filter=True):
    from niwqg_b200 import CoupledModel, YBJModel
    from oracle import niwqg_oracle as orc
    kw, U0, k0 = lamb_params(nx, use_filter, 10 ** 9, nsteps)
    kw["twrite"] = 10 ** 9
    cls = {"coupled": CoupledModel, "ybj": YBJModel}[model].Model
    m = cls(**kw)
    o = orc.NIWQGOracle(model=model, **kw)
    np.random.seed(7)
    q = orc.mcwilliams1984(o, k0=k0, E=U0 ** 2 / 2)          # random red spectrum (BASELINE config 3)
    phi = (np.ones_like(q) + 1j) * (2 * U0) / np.sqrt(2)
    for mdl in (m, o):
        mdl.set_q(q); mdl.set_phi(phi)
    return m, o


def test_coupled_1024_random_spectrum_matches_oracle():
    m, o = _pair("coupled", 1024, 3)
    for _ in range(3):
        m._step_forward(); o.step_forward()
    assert rel_l2(m.q, o.q) < 1e-10 and rel_l2(m.phi, o.phi) < 1e-10
    for k in ("Ke", "Pw", "Kw"):
        assert abs(getattr(m, k) - getattr(o, k)) <= 1e-10 * max(abs(getattr(o, k)), 1e-3 * abs(o.Kw)), k


def test_ybj_2048_random_spectrum_matches_oracle():
    m, o = _pair("ybj", 2048, 3)
    for _ in range(3):
        m._step_forward(); o.step_forward()
    assert rel_l2(m.phi, o.phi) < 1e-10


def test_8192_linear_propagator_is_exact():
    """niwqg/tests/test_diffusion.py at the target grid: with nu4 only and a plane-wave q, phi = 0, ETDRK4 must
    reproduce qh0 * exp(-nu4 wv4 t) (tables, transforms through the cluster kernels, stage kernels at full size)."""
    from niwqg_b200 import CoupledModel
    N = 8192
    kx, ky = 37, 4001            # one mode near the grid scale in y: strong, exactly known decay
    L = 5e5
    k, l = 2 * np.pi * kx / L, 2 * np.pi * ky / L
    nu4 = 2.0 / ((k * k + l * l) ** 2 * 3 * 2000.)          # e^-2 over the three steps
    m = CoupledModel.Model(nx=N, L=L, use_filter=False, nu4=nu4, nu4w=0., nu=0., nuw=0., dt=2000., tmax=3 * 2000. - 1.,
                           twrite=10 ** 9, tdiags=10 ** 9)
    xs = (np.arange(N) + 0.5) / N * m.L
    A = 1e-10                    # small amplitude: the (analytically vanishing) Jacobian's rounding noise stays negligible
    qi = A * np.sin(k * xs[None, :] + l * xs[:, None])
    m.set_q(qi); m.set_phi(np.zeros((N, N), complex))
    m.run()
    assert m.tc == 3
    qh = m.qh
    decay = np.exp(-nu4 * (k * k + l * l) ** 2 * 3 * 2000.)
    # qi = Im exp(i(kx+ly)): modes (ky,kx) and (-ky,-kx) with amplitude N^2/2 and the half-cell phase
    amp = abs(qh[ky, kx]) / (N * N / 2) / A
    assert abs(amp - decay) < 1e-10 * decay, (amp, decay)
    assert abs(abs(qh[N - ky, N - kx]) / (N * N / 2) / A - decay) < 1e-10 * decay
    qh[ky, kx] = 0; qh[N - ky, N - kx] = 0
    assert np.abs(qh).max() < 1e-5 * (N * N / 2) * A * decay
    # and the physical field is the decayed plane wave
    assert rel_l2(m.q, decay * qi) < 1e-6     # the Jacobian rounding noise of all other modes (the reference shows the same)


def test_8192_step_conserves_and_budgets_close():
    """One Coupled step at 8192^2 from the bench initial condition: finite, the integrated budgets follow the diagnosed
    energies, and a second model stepped from the same state agrees bit for bit (deterministic reductions)."""
    from niwqg_b200 import CoupledModel, InitialConditions as ic
    N = 8192
    kw, U0, k0 = lamb_params(N, True, 1, 2)
    kw["twrite"] = 10 ** 9
    m = CoupledModel.Model(**kw)
    q = ic.LambDipole(m, U=U0, R=2 * np.pi / k0)
    phi = (np.ones_like(q) + 1j) * (2 * U0) / np.sqrt(2)
    m.set_q(q); m.set_phi(phi)
    m._step_forward(); m._step_forward()
    d = m.diagnostics
    ke, Ke = d["ke_qg"]["value"], d["Ke"]["value"]
    assert np.all(np.isfinite(ke)) and np.all(np.isfinite(d["ke_niw"]["value"]))
    assert abs(ke[-1] - Ke[-1]) < 1e-6 * ke[0]
    assert abs(d["ke_niw"]["value"][-1] - d["Kw"]["value"][-1]) < 1e-6 * d["ke_niw"]["value"][0]
    q1 = m.q
    del m
    m2 = CoupledModel.Model(**kw)
    m2.set_q(q); m2.set_phi(phi)
    m2._step_forward(); m2._step_forward()
    assert np.array_equal(q1, m2.q)


@pytest.mark.parametrize("N", [2048, 4096, 8192])
def test_cluster_transforms_are_bitwise_repeatable(N):
    """The cluster kernels exchange through distributed shared memory behind barriers; a missing barrier would show up
    as run-to-run differences (compute-sanitizer's racecheck is not available on the GPU pool).  Same input, several
    back-to-back transforms (so different CTA placements and timings), bit-identical outputs, forward and inverse."""
    from niwqg_b200 import _native as nat
    h = nat.Handle(model=nat.MODEL_YBJ, nx=N, batch=1, device=0, L=5e5, dt=1e4, f=1e-4, N=0.01, m=0.025, nu=20., nuw=50.)
    rng = np.random.RandomState(11)
    x = rng.randn(N, N) + 1j * rng.randn(N, N)
    X0 = h.fft2(x, nat.FFT_C2C_FWD)
    x0 = h.fft2(X0, nat.FFT_C2C_INV)
    for _ in range(3):
        assert np.array_equal(h.fft2(x, nat.FFT_C2C_FWD), X0)
        assert np.array_equal(h.fft2(X0, nat.FFT_C2C_INV), x0)
    assert rel_l2(x0, x) < 5e-15
    h.close()
