"""save_to_disk=True end to end with the reference's HDF5 layout (niwqg/Saving.py:38-101; call sites Kernel.py:147-148,
:194-195, :202-203, :217): dataset names, shapes and cadence for the kernel family and for QGModel.  h5py is not
installed in this image, so a recording stand-in (tests/fake_h5py.py) takes its place."""
import os
import sys

import numpy as np
import pytest

from cases import lamb_params

pytestmark = pytest.mark.gpu


@pytest.fixture
def h5(monkeypatch):
    import fake_h5py
    fake_h5py.WRITTEN.clear()
    monkeypatch.setitem(sys.modules, "h5py", fake_h5py)
    import logging
    logging.disable(logging.CRITICAL)
    return fake_h5py


@pytest.mark.parametrize("model", ["coupled", "qg"])
def test_saving_layout_and_cadence(model, h5, tmp_path):
    from niwqg_b200 import CoupledModel, QGModel
    from oracle import niwqg_oracle as orc
    qg = model == "qg"
    nx, nsteps, tsnaps = 64, 7, 3
    kw, U0, k0 = lamb_params(nx, True, 2, nsteps, qg=qg)
    path = str(tmp_path / "out")
    m = (QGModel if qg else CoupledModel).Model(save_to_disk=True, tsave_snapshots=tsnaps, path=path, **kw)
    # setup.h5 is written by the constructor (Kernel.py:147-148)
    setup = h5.WRITTEN[path + "/setup.h5"]
    assert sorted(setup) == ["grid/k", "grid/l", "grid/nx", "grid/wv", "grid/x", "grid/y"]
    assert int(setup["grid/nx"]) == nx and setup["grid/x"].shape == (nx, nx) and setup["grid/y"].shape == (nx, nx)
    assert setup["grid/wv"].shape == ((nx, nx // 2 + 1) if qg else (nx, nx))
    assert setup["grid/k"].shape == ((nx // 2 + 1,) if qg else (nx,)) and setup["grid/l"].shape == (nx,)
    assert np.array_equal(setup["grid/x"], m.x) and np.array_equal(setup["grid/wv"], m.wv)
    q = orc.lamb_dipole(m, U=U0, R=2 * np.pi / k0)
    m.set_q(q)
    if not qg:
        m.set_phi((np.ones_like(q) + 1j) * (2 * U0) / np.sqrt(2))
    m.run()
    assert m.tc == nsteps
    # run() writes the initial state (tc = 0), then _step_forward every tsnaps steps (Kernel.py:194-195, :217)
    snaps = sorted(p for p in h5.WRITTEN if "/snapshots/" in p)
    want_tc = [0] + [tc for tc in range(1, nsteps + 1) if tc % tsnaps == 0]
    t, want_names = 0.0, []
    tcs = {}
    for tc in range(nsteps + 1):
        tcs[tc] = t
        t += m.dt
    want_names = [path + "/snapshots/{:015.0f}.h5".format(tcs[tc]) for tc in want_tc]
    assert snaps == sorted(want_names)
    fields = ["t", "q"] if qg else ["t", "q", "phi"]              # QGModel.py:196-199 / Kernel.py:195
    for p in snaps:
        d = h5.WRITTEN[p]
        assert sorted(d) == sorted(fields), (p, sorted(d))
        assert d["q"].shape == (nx, nx) and d["q"].dtype == np.float64
        if not qg:
            assert d["phi"].shape == (nx, nx) and d["phi"].dtype == np.complex128
    last = h5.WRITTEN[want_names[-1]]
    assert float(last["t"]) == tcs[want_tc[-1]]
    first = h5.WRITTEN[want_names[0]]
    assert np.allclose(first["q"], q, rtol=0, atol=1e-12 * np.abs(q).max())     # the seeded state, through fft/ifft
    # diagnostics.h5: one dataset per registered diagnostic, the accumulated series (Saving.py:88-101)
    diag = h5.WRITTEN[path + "/diagnostics.h5"]
    assert sorted(diag) == sorted(m.diagnostics.keys())
    ntick = len([tc for tc in range(nsteps) if tc % m.tdiags == 0])
    assert np.atleast_1d(diag["time"]).shape == np.atleast_1d(m.diagnostics["time"]["value"]).shape
    assert np.atleast_1d(diag["ke_qg"]).size in (ntick, ntick + 1, ntick - 1)
    assert os.path.exists(path + "/diagnostics.h5")


def test_saving_refuses_slab_runs(h5, tmp_path):
    from niwqg_b200 import CoupledModel
    with pytest.raises(NotImplementedError):
        CoupledModel.Model(nx=64, save_to_disk=True, path=str(tmp_path / "o"), rank=0, nranks=2, nccl_id=b"\0" * 128)
