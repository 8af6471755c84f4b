"""`unidistill` mirror: the reference's module paths for the BEV hot path, backed by
libunidistill_b200.so (hand-written sm_100a kernels).  See DESIGN.md / INTEGRATION.md."""
from ._overlay import extend as _extend

_extend(__path__, __name__, globals())
__version__ = "0.1.0+b200"
