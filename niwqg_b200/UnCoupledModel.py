"""Young & Ben Jelloul uncoupled NIW model over an evolving QG flow (niwqg/UnCoupledModel.py)."""
from . import Kernel
from . import _native as nat


class Model(Kernel.Kernel):
    _model_id = nat.MODEL_UNCOUPLED

    def __init__(self, **kwargs):
        self.model = " Uncoupled Model"
        super(Model, self).__init__(**kwargs)
