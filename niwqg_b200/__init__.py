"""niwqg_b200: B200-native implementation of niwqg's pseudo-spectral ETDRK4 hot path.

Drop-in for the reference package layout::

    import niwqg_b200 as niwqg
    from niwqg_b200 import CoupledModel, QGModel, InitialConditions as ic
    m = CoupledModel.Model(nx=512, ...); m.set_q(q); m.set_phi(phi); m.run()
"""
__version__ = '0.1beta-b200'

from . import Diagnostics
from . import InitialConditions
from . import Saving
