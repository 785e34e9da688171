"""Quasilinear Xie & Vanneste model (niwqg/QLModel.py).

The reference file does not run as shipped (no ``self.model``, undefined
``jacobian_phic_phi``, ``q`` never refreshed - SURVEY.md F8).  This class is the
repaired reading used as the oracle: CoupledModel's inversion with the wave
advected by the vortex flow ``psi_q = -wv2i*qh`` only (niwqg/QLModel.py:65-67).
"""
from . import CoupledModel
from . import _native as nat


class Model(CoupledModel.Model):
    _model_id = nat.MODEL_QL

    def __init__(self, **kwargs):
        super(Model, self).__init__(**kwargs)
        self.model = " Quasilinear Model"
