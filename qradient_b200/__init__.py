"""qradient_b200 -- the state-vector simulator and adjoint-gradient hot path of
frederikwilde/qradient, rebuilt from scratch on hand-written sm_100a CUDA kernels.

Drop-in surface (same names, arguments and error behaviour as the reference):

    from qradient_b200.circuit_logic import McClean, Qaoa
    from qradient_b200.physical_components import State, Gates, Observable
    from qradient_b200.optimization_problems import MaxCut
    from qradient_b200.optimization import McCleanOpt, QaoaOpt        # host optimiser loops around grad_run

All numerics run on the GPU through the C ABI in include/qradient_b200.h; there is no CPU path.
"""
from . import _lib  # noqa: F401
from . import physical_components, circuit_logic, optimization_problems, optimization  # noqa: F401
from .circuit_logic import McClean, Qaoa  # noqa: F401
from .physical_components import State, Gates, Observable  # noqa: F401

__version__ = "0.1.0"
