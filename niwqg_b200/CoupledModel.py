"""Xie & Vanneste (2015) coupled NIW-QG model (niwqg/CoupledModel.py) on the CUDA backend."""
from . import Kernel
from . import _native as nat
from .Diagnostics import add_diagnostic


class Model(Kernel.Kernel):
    _model_id = nat.MODEL_COUPLED
    qw = Kernel._DeviceField("QW")
    qwh = Kernel._DeviceField("QWH")

    def __init__(self, **kwargs):
        self.model = " Coupled Model"
        super(Model, self).__init__(**kwargs)

    def jacobian_phic_phi(self):
        """niwqg/CoupledModel.py:59-73 (refreshes phix, phiy on the device)."""
        return self._h.jacobian(nat.JAC_PHIC_PHI)

    def _initialize_class_diagnostics(self):
        """niwqg/CoupledModel.py:115-136."""
        S = nat.S
        for name, desc, slot in (('ke_qg_q', 'Quasigeostrophic Kinetic Energy, q-flow', "KE_QG_Q"),
                                 ('ke_qg_w', 'Quasigeostrophic Kinetic Energy, w-flow', "KE_QG_W"),
                                 ('ke_qg_qw', 'Quasigeostrophic Kinetic Energy, cross-term q-w', "KE_QG_QW")):
            add_diagnostic(self, name, description=desc, units=r'm^2 s^{-2}', types='scalar',
                           function=(lambda self, slot=slot: self._diag[S[slot]]))

    def _calc_class_derived_fields(self):
        """niwqg/CoupledModel.py:138-143."""
        S = nat.S
        self.ke_qg_q, self.ke_qg_w, self.ke_qg_qw = (self._diag[S["KE_QG_Q"]], self._diag[S["KE_QG_W"]],
                                                     self._diag[S["KE_QG_QW"]])
