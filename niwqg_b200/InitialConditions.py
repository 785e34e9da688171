"""Initial-condition generators with the reference's signatures
(niwqg/InitialConditions.py).  By default they are host-side numpy/scipy (the arrays the
numpy oracle is seeded with, bit for bit); the model protocol they use (``model.x, y, wv,
wv2, fft, ifft, spec_var``) is served by the CUDA backend.

``on_device=True`` generates the field on the device instead (csrc/kernels_ic.cuh, C ABI
``niwqg_ic``) and seeds the model with it directly - the equivalent of
``model.set_q(LambDipole(model, ...))`` without any whole-grid host array (at 8192^2 the
host path needs several GiB and, for the random spectra, four host round trips of the FFT
seam).  The call then returns None.  Random spectra take their phases from the global numpy
RNG like the reference unless ``seed`` is given, in which case they come from a Philox
counter-based generator on the device.
"""
import numpy as np


def _device(model, kind, params, rand01=None):
    model._h.ic(kind, params, rand01)
    if kind in ("LambDipole", "McWilliams1984", "Danioux2015") and hasattr(model, "Ke") and not getattr(model, "_is_qg", False):
        model.ke = None                # Kernel.set_q side effect (niwqg/Kernel.py:535): served from the device on first read
    return None


def _device_spectrum(model, kind, k0, E, seed):
    if seed is None:
        nhx, nhy = model.nx, model.nx
        return _device(model, kind, [k0, E, 0], np.random.rand(nhx, nhy))      # same draw as the host generator
    return _device(model, kind, [k0, E, float(seed)])


def _random_red_spectrum(model, ckappa, E):
    # niwqg/InitialConditions.py:34-41 / :68-75: random phases from the global numpy RNG
    nhx, nhy = model.wv2.shape
    phase = np.random.rand(nhx, nhy) * 2 * np.pi
    ph = ckappa * np.cos(phase) + 1j * ckappa * np.sin(phase)
    ph = model.fft(model.ifft(ph).real)
    Eaux = 0.5 * model.spec_var(model.wv * ph)
    pih = np.sqrt(E / Eaux) * ph
    return model.ifft(-model.wv2 * pih).real


def McWilliams1984(model, k0=6, E=0.5, on_device=False, seed=None):
    """Random vorticity with the red spectrum of McWilliams (1984).  niwqg/InitialConditions.py:4-41."""
    if on_device:
        return _device_spectrum(model, "McWilliams1984", k0, E, seed)
    ckappa = np.zeros_like(model.wv2)
    fk = model.wv != 0
    ckappa[fk] = np.sqrt(model.wv2[fk] * (1. + (model.wv2[fk] / k0 ** 2) ** 2)) ** -1
    return _random_red_spectrum(model, ckappa, E)


def Danioux2015(model, k0=6, E=0.5, on_device=False, seed=None):
    """Single-wavenumber-band random vorticity.  niwqg/InitialConditions.py:43-75."""
    if on_device:
        return _device_spectrum(model, "Danioux2015", k0, E, seed)
    ckappa = np.zeros_like(model.wv2)
    fk = model.wv != 0
    ckappa[fk] = np.sqrt(model.wv[fk] * np.exp(-(model.wv2[fk] / k0 ** 2)))
    return _random_red_spectrum(model, ckappa, E)


def LambDipole(model, U=.01, R=1., on_device=False):
    """Lamb dipole vorticity.  niwqg/InitialConditions.py:77-114 (the reference's O(N^2)
    Python loop only guards the division at r == 0; vectorised here)."""
    if on_device:
        return _device(model, "LambDipole", [U, R])
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


def WavePacket(model, k=10, l=0, R=1, x0=0., y0=0., on_device=False):
    """Gaussian wave packet.  niwqg/InitialConditions.py:117-145."""
    if on_device:
        return _device(model, "WavePacket", [k, l, R, x0, y0])
    x, y = model.x, model.y
    r = np.sqrt((x - x0) ** 2 + (y - y0) ** 2)
    phi = np.exp(1j * (k * (x - x0) + l * (y - y0)))
    phi *= np.exp(-((r / R) ** 2))
    return phi


def UniformWave(model, phi0=0.2 * (1 + 1j) / np.sqrt(2), on_device=True):
    """Uniform near-inertial wave phi = phi0 (examples/LambDipole.py:52), seeded on the device."""
    if on_device:
        return _device(model, "Uniform", [complex(phi0).real, complex(phi0).imag])
    return np.ones((model.ny, model.nx)) * complex(phi0)


def PlaneWave(model, k=10, l=0, phase=0., on_device=False):
    """Plane wave; as in the reference the ``phase`` is added outside ``1j*`` and so
    scales the amplitude (niwqg/InitialConditions.py:147-169)."""
    if on_device:
        return _device(model, "PlaneWave", [k, l, phase])
    return np.exp(1j * (k * model.x + l * model.y) + phase)
