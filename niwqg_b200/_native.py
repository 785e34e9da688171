"""ctypes binding of libniwqg_b200.so (C ABI declared in include/niwqg_b200.h).

There is no CPU fallback: if the shared library has not been built
(``python -c "import __graft_entry__ as g; g.build()"``) or no CUDA device is
present, model construction raises RuntimeError.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libniwqg_b200.so")

MODEL_QG, MODEL_COUPLED, MODEL_UNCOUPLED, MODEL_YBJ, MODEL_QL = range(5)

# enum niwqg_field
FIELDS = ["Q", "QH", "PH", "P", "PHI", "PHIH", "PHIX", "PHIY", "LAPPHI", "U", "V", "QW", "QPSI", "QWH", "C", "CH",
          "FILTR", "EXPCH", "EXPCH_H", "QHCOEF", "F0", "FAB", "FC", "EXPCHW", "EXPCH_HW", "QHWCOEF", "F0W", "FABW",
          "FCW", "EXPCHC", "EXPCH_HC", "QHCCOEF", "F0C", "FABC", "FCC"]
F = {n: i for i, n in enumerate(FIELDS)}
REAL_FIELDS = {"Q", "P", "U", "V", "QW", "QPSI", "C", "FILTR"}
PHYS_CPLX_FIELDS = {"PHI", "PHIX", "PHIY", "LAPPHI"}
TABLE_FIELDS = set(FIELDS[F["FILTR"]:])

# enum niwqg_scalar
SCALARS = ["KE", "PW", "KW", "GAMMA1", "GAMMA2", "XI1", "XI2", "PI", "KE_QG", "ENS", "KE_NIW", "CKE_NIW", "IKE_NIW",
           "PE_NIW", "CONC", "SKEW", "EP_PHI", "EP_PSI", "CHI_Q", "CHI_PHI", "KE_QG_Q", "KE_QG_W", "KE_QG_QW", "CFL",
           "CVAR", "C2", "GRADC2", "GAMMA_C", "EP_C", "CHI_C"]
S = {n: i for i, n in enumerate(SCALARS)}
S_COUNT = len(SCALARS)

FFT_C2C_FWD, FFT_C2C_INV, FFT_R2C, FFT_C2R, FFT_R2C_FULL = range(5)
JAC_PSI_Q, JAC_PHIC_PHI, JAC_PSI_PHI = range(3)

EXPORTS = ["niwqg_create", "niwqg_destroy", "niwqg_last_error", "niwqg_set_q", "niwqg_set_phi", "niwqg_set_c",
           "niwqg_step", "niwqg_diagnostics", "niwqg_status", "niwqg_get_scalars", "niwqg_get_field",
           "niwqg_field_bytes", "niwqg_fft2", "niwqg_jacobian", "niwqg_sync", "niwqg_time_steps",
           "niwqg_launch_count", "niwqg_stream", "niwqg_profile"]


class Params(C.Structure):
    _fields_ = [("model", C.c_int), ("nx", C.c_int), ("batch", C.c_int), ("device", C.c_int),
                ("L", C.c_double), ("dt", C.c_double), ("U", C.c_double), ("f", C.c_double), ("N", C.c_double),
                ("m", C.c_double), ("nu", C.c_double), ("nu4", C.c_double), ("mu", C.c_double),
                ("nuw", C.c_double), ("nu4w", C.c_double), ("muw", C.c_double), ("beta", C.c_double),
                ("use_filter", C.c_int), ("dealias", C.c_int), ("passive_scalar", C.c_int),
                ("nu4c", C.c_double), ("nuc", C.c_double), ("muc", C.c_double)]


_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("niwqg_b200: %s not built (run __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, ip, dp = C.c_void_p, C.c_int, C.POINTER(C.c_double)
    lib.niwqg_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    lib.niwqg_destroy.argtypes = [vp]
    lib.niwqg_last_error.argtypes = [vp]
    lib.niwqg_last_error.restype = C.c_char_p
    for n in ("niwqg_set_q", "niwqg_set_phi", "niwqg_set_c"):
        getattr(lib, n).argtypes = [vp, vp, ip]
    lib.niwqg_step.argtypes = [vp, ip]
    for n in ("niwqg_diagnostics", "niwqg_status", "niwqg_get_scalars"):
        getattr(lib, n).argtypes = [vp, vp]
    lib.niwqg_get_field.argtypes = [vp, ip, ip, vp, C.c_size_t, ip]
    lib.niwqg_field_bytes.argtypes = [vp, ip]
    lib.niwqg_field_bytes.restype = C.c_size_t
    lib.niwqg_fft2.argtypes = [vp, vp, vp, ip]
    lib.niwqg_jacobian.argtypes = [vp, ip, vp]
    lib.niwqg_sync.argtypes = [vp]
    lib.niwqg_time_steps.argtypes = [vp, ip, C.POINTER(C.c_float)]
    lib.niwqg_launch_count.argtypes = [vp]
    lib.niwqg_launch_count.restype = C.c_longlong
    lib.niwqg_profile.argtypes = [vp, ip, vp, vp]
    lib.niwqg_stream.argtypes = [vp]
    lib.niwqg_stream.restype = vp
    _lib = lib
    return lib


class Handle(object):
    """Owns one niwqg_handle.  Every method raises RuntimeError on a non-zero return."""

    def __init__(self, **kw):
        self.lib = load()
        p = Params()
        for k, v in kw.items():
            setattr(p, k, v)
        self.params = p
        self.h = C.c_void_p()
        rc = self.lib.niwqg_create(C.byref(p), C.byref(self.h))
        if rc != 0:
            msg = self.lib.niwqg_last_error(None)
            self.h = None
            raise RuntimeError("niwqg_create failed (%d): %s" % (rc, msg.decode() if msg else "?"))
        self.N, self.B = p.nx, p.batch
        self.nk = p.nx // 2 + 1 if p.model == MODEL_QG else p.nx

    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.niwqg_last_error(self.h)
            raise RuntimeError("niwqg_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))

    def close(self):
        if getattr(self, "h", None):
            self.lib.niwqg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- seeding -----------------------------------------------------------
    def _host(self, a, dtype):
        a = np.ascontiguousarray(a, dtype=dtype)
        want = self.B * self.N * self.N
        if a.size != want:
            if a.size == self.N * self.N and self.B > 1:
                a = np.ascontiguousarray(np.broadcast_to(a.reshape(1, self.N, self.N), (self.B, self.N, self.N)))
            else:
                raise ValueError("expected %d values, got %d" % (want, a.size))
        return a

    def set_q(self, q):
        a = self._host(q, np.float64)
        self._ck(self.lib.niwqg_set_q(self.h, a.ctypes.data, 0))
        self.sync()

    def set_phi(self, phi):
        a = self._host(phi, np.complex128)
        self._ck(self.lib.niwqg_set_phi(self.h, a.ctypes.data, 0))
        self.sync()

    def set_c(self, c):
        a = self._host(c, np.float64)
        self._ck(self.lib.niwqg_set_c(self.h, a.ctypes.data, 0))
        self.sync()

    def set_q_device(self, ptr):
        self._ck(self.lib.niwqg_set_q(self.h, C.c_void_p(ptr), 1))

    def set_phi_device(self, ptr):
        self._ck(self.lib.niwqg_set_phi(self.h, C.c_void_p(ptr), 1))

    # -- stepping ----------------------------------------------------------
    def step(self, n=1):
        self._ck(self.lib.niwqg_step(self.h, int(n)))

    def time_steps(self, n):
        ms = C.c_float()
        self._ck(self.lib.niwqg_time_steps(self.h, int(n), C.byref(ms)))
        return ms.value

    def sync(self):
        self._ck(self.lib.niwqg_sync(self.h))

    PROFILE_KINDS = ("fft_row", "fft_col", "phys", "spec", "small")

    def profile(self, enable):
        """Switch per-kernel-kind event timing on/off; returns {kind: (total_ms, launches)} recorded so far."""
        ms = np.zeros(5)
        cnt = np.zeros(5, np.int64)
        self._ck(self.lib.niwqg_profile(self.h, int(bool(enable)), ms.ctypes.data, cnt.ctypes.data))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.PROFILE_KINDS)}

    def field_into(self, name, out, member=0):
        """Copy one member's field into a caller-provided (e.g. pinned) host array."""
        self._ck(self.lib.niwqg_get_field(self.h, F[name], member, out.ctypes.data, out.nbytes, 0))
        return out

    def launch_count(self):
        return int(self.lib.niwqg_launch_count(self.h))

    # -- reads -------------------------------------------------------------
    def scalars(self, which="scalars"):
        out = np.zeros((self.B, S_COUNT))
        fn = {"scalars": self.lib.niwqg_get_scalars, "diagnostics": self.lib.niwqg_diagnostics}[which]
        self._ck(fn(self.h, out.ctypes.data))
        return out

    def status(self):
        out = np.zeros((self.B, 4))
        self._ck(self.lib.niwqg_status(self.h, out.ctypes.data))
        return out

    def field(self, name, member=None):
        fid = F[name]
        N, nk = self.N, self.nk
        if name in REAL_FIELDS:
            shape, dt = ((N, nk) if name == "FILTR" else (N, N)), np.float64
        elif name in PHYS_CPLX_FIELDS:
            shape, dt = (N, N), np.complex128
        else:
            shape, dt = (N, nk), np.complex128
        members = [0] if (name in TABLE_FIELDS) else (range(self.B) if member is None else [member])
        outs = []
        for mm in members:
            a = np.empty(shape, dt)
            self._ck(self.lib.niwqg_get_field(self.h, fid, mm, a.ctypes.data, a.nbytes, 0))
            outs.append(a)
        if len(outs) == 1:
            return outs[0]
        return np.stack(outs)

    def field_to_device(self, name, member, ptr, nbytes):
        self._ck(self.lib.niwqg_get_field(self.h, F[name], member, C.c_void_p(ptr), nbytes, 1))

    def fft2(self, x, kind):
        N, nh = self.N, self.N // 2 + 1
        if kind in (FFT_R2C, FFT_R2C_FULL):
            a = np.ascontiguousarray(x, np.float64)
            out = np.empty((N, nh) if kind == FFT_R2C else (N, N), np.complex128)
        elif kind == FFT_C2R:
            a = np.ascontiguousarray(x, np.complex128)
            if a.shape != (N, nh):
                raise ValueError("irfft2 input must be (%d,%d)" % (N, nh))
            out = np.empty((N, N), np.float64)
        else:
            a = np.ascontiguousarray(x, np.complex128)
            out = np.empty((N, N), np.complex128)
        if kind != FFT_C2R and a.shape != (N, N):
            raise ValueError("fft input must be (%d,%d)" % (N, N))
        self._ck(self.lib.niwqg_fft2(self.h, a.ctypes.data, out.ctypes.data, kind))
        return out

    def jacobian(self, which):
        out = np.empty((self.N, self.nk), np.complex128)
        self._ck(self.lib.niwqg_jacobian(self.h, which, out.ctypes.data))
        return out
