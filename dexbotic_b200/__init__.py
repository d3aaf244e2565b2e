"""dexbotic_b200 — B200 (sm_100a) backend for Dexbotic's VLA training hot path.

Layout:
  csrc/        hand-written CUDA (tcgen05 GEMM, HBM-bound operators) + the C-ABI (include/*.h)
  _lib.py      ctypes binding of libdexbotic_b200.so
  ops.py       tensor-level wrappers (one call = one kernel launch)
  ...          host-side mirror of dexbotic/model (same class names, forward signatures, state-dict keys)
"""
__version__ = "0.1.0"
