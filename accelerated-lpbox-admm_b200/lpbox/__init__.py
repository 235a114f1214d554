"""`lpbox` -- host-side mirror of the reference's Cython module of the same name.

Reference surface being mirrored (same names, argument meaning and return values):
  * LP:  `LinerProgramming/LinearProgramming/cython_solver/lpbox.pyx:7-76`  -> :class:`PyLPboxADMMsolver`
  * batched entry points (new; the reference solves one instance per object) -> :class:`LPBatch`
All compute happens in hand-written sm_100a CUDA behind the C ABI of `include/lpbox_b200.h`.
"""
from .lp import LPBatch, PyLPboxADMMsolver, gen_auctions, read_instance  # noqa: F401
from . import _capi  # noqa: F401
from .l2f import solve_l2f  # noqa: F401
from . import sparse_attack  # noqa: F401
from .seg import PySegLPboxADMMsolver, SegBatch, build_graph  # noqa: F401
