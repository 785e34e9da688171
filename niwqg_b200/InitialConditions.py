"""Initial-condition generators with the reference's signatures
(niwqg/InitialConditions.py).  One-off host-side setup (numpy/scipy); the model
protocol they use (``model.x, y, wv, wv2, fft, ifft, spec_var``) is served by the
CUDA backend.
"""
import numpy as np


def _random_red_spectrum(model, ckappa, E):
    # niwqg/InitialConditions.py:34-41 / :68-75: random phases from the global numpy RNG
    nhx, nhy = model.wv2.shape
    phase = np.random.rand(nhx, nhy) * 2 * np.pi
    ph = ckappa * np.cos(phase) + 1j * ckappa * np.sin(phase)
    ph = model.fft(model.ifft(ph).real)
    Eaux = 0.5 * model.spec_var(model.wv * ph)
    pih = np.sqrt(E / Eaux) * ph
    return model.ifft(-model.wv2 * pih).real


def McWilliams1984(model, k0=6, E=0.5):
    """Random vorticity with the red spectrum of McWilliams (1984).  niwqg/InitialConditions.py:4-41."""
    ckappa = np.zeros_like(model.wv2)
    fk = model.wv != 0
    ckappa[fk] = np.sqrt(model.wv2[fk] * (1. + (model.wv2[fk] / k0 ** 2) ** 2)) ** -1
    return _random_red_spectrum(model, ckappa, E)


def Danioux2015(model, k0=6, E=0.5):
    """Single-wavenumber-band random vorticity.  niwqg/InitialConditions.py:43-75."""
    ckappa = np.zeros_like(model.wv2)
    fk = model.wv != 0
    ckappa[fk] = np.sqrt(model.wv[fk] * np.exp(-(model.wv2[fk] / k0 ** 2)))
    return _random_red_spectrum(model, ckappa, E)


def LambDipole(model, U=.01, R=1.):
    """Lamb dipole vorticity.  niwqg/InitialConditions.py:77-114 (the reference's O(N^2)
    Python loop only guards the division at r == 0; vectorised here)."""
    from scipy import special
    N = model.nx
    x, y = model.x, model.y
    x0, y0 = x[N // 2, N // 2], y[N // 2, N // 2]
    r = np.sqrt((x - x0) ** 2 + (y - y0) ** 2)
    s = np.zeros_like(r)
    nz = r != 0.
    s[nz] = (y[nz] - y0) / r[nz]
    lam = (3.8317) / R
    C = -(2. * U * lam) / (special.j0(lam * R))
    q = np.zeros_like(r)
    inside = r <= R
    q[inside] = C * special.j1(lam * r[inside]) * s[inside]
    return q


def WavePacket(model, k=10, l=0, R=1, x0=0., y0=0.):
    """Gaussian wave packet.  niwqg/InitialConditions.py:117-145."""
    x, y = model.x, model.y
    r = np.sqrt((x - x0) ** 2 + (y - y0) ** 2)
    phi = np.exp(1j * (k * (x - x0) + l * (y - y0)))
    phi *= np.exp(-((r / R) ** 2))
    return phi


def PlaneWave(model, k=10, l=0, phase=0.):
    """Plane wave; as in the reference the ``phase`` is added outside ``1j*`` and so
    scales the amplitude (niwqg/InitialConditions.py:147-169)."""
    return np.exp(1j * (k * model.x + l * model.y) + phase)
