"""Young & Ben Jelloul NIW model over a STEADY QG flow (niwqg/YBJModel.py)."""
from . import Kernel
from . import _native as nat


class Model(Kernel.Kernel):
    _model_id = nat.MODEL_YBJ

    def __init__(self, **kwargs):
        self.model = " YBJ Model (Steady QG flow)"
        super(Model, self).__init__(**kwargs)
