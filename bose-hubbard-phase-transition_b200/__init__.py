"""bose-hubbard-phase-transition_b200 -- B200-native exact-diagonalisation hot path.

The product is libbh_b200.so (hand-written sm_100a CUDA kernels behind the C ABI of include/bh_b200.h)
and the C++ host layer under host/.  This Python package only binds the C ABI for the tests and
bench.py (see capi.py).  The directory name is not a Python identifier: load it with
`__graft_entry__.load_package()`.
"""
from . import capi  # noqa: F401
from .capi import Context, BhError  # noqa: F401
