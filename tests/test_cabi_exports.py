"""CPU test: the C-ABI shared library loads here (no GPU needed) and exports every symbol that
include/*.h declares; calling a compute entry point without a device fails loudly, never silently."""
import ctypes

import pytest

from dexbotic_b200 import _lib


def test_library_built_and_loads():
    assert _lib.lib_available(), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    lib = _lib.load()
    assert lib.b200_version() >= 100


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    assert len(_lib.EXPORTED) >= 30
    for name in _lib.EXPORTED:
        assert hasattr(lib, name), f"include/*.h declares {name} but the library does not export it"


def test_gemm_args_struct_matches_header():
    # field order/size sanity: the ctypes mirror must have the same size on both sides of the ABI
    assert ctypes.sizeof(_lib.GemmArgs) % 8 == 0
    assert _lib.GemmArgs.a.offset == 0 and _lib.GemmArgs.d.offset == 24 and _lib.GemmArgs.m.offset == 48


def test_no_cpu_fallback():
    import torch
    from dexbotic_b200 import ops
    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(8, 8), torch.zeros(8, 8))
    with pytest.raises(RuntimeError):
        ops.rmsnorm_fwd(torch.zeros(4, 8), torch.ones(8), 1e-6)
