"""The C-ABI library loads without a GPU and exports every symbol include/lavt_b200.h declares (no compute calls)."""
import ctypes
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "lavt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lavt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from lavt_rs_b200.build import build_library
    path = build_library()
    lib = ctypes.CDLL(path)
    syms = _header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in lavt_b200.h but not exported"


def test_ctypes_binding_covers_header():
    from lavt_rs_b200 import _cabi
    assert sorted(_cabi.EXPORTS) == _header_symbols()
    lib = _cabi.lib()
    assert lib.lavt_abi_version() == _cabi.ABI_VERSION


def test_struct_layouts_match_header():
    from lavt_rs_b200 import _cabi
    assert ctypes.sizeof(_cabi.WinGeom) == 17 * 4
    assert ctypes.sizeof(_cabi.Epilogue) == 9 * 8 + 6 * 4      # 6 data pointers + win + rscale + out_pre pointers, act/ldm/ldo/mul_act/rscale_rows/pre_mode
    assert _cabi.Epilogue.mul_act.offset == 60 and _cabi.Epilogue.out_pre.offset == 88
    assert _cabi.Epilogue.win.offset == 64 and _cabi.Epilogue.rscale.offset == 72 and _cabi.Epilogue.rscale_rows.offset == 80
    from lavt_rs_b200.optim import _Tensor
    assert ctypes.sizeof(_Tensor) == 56                         # lavt_adamw_tensor_t: 5 pointers, int64 n, two floats


def test_no_cpu_fallback():
    import pytest
    import torch
    from lavt_rs_b200 import _cabi
    a = torch.zeros(128, 64, dtype=torch.bfloat16)
    with pytest.raises(_cabi.LavtError):
        _cabi.gemm_bf16(a, a, out_bf16=torch.zeros(128, 128, dtype=torch.bfloat16))
