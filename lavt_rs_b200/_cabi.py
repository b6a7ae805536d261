"""ctypes binding of ``include/lavt_b200.h`` (the C-ABI of the sm_100a kernels).

The product path has NO CPU fallback: if the shared library is missing or the device is not a
B200, every call raises.  PyTorch tensors are used only as device-memory containers; the library
sees raw pointers and the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "_lib", "liblavt_b200.so")

ACT_NONE, ACT_GELU, ACT_RELU, ACT_TANH = 0, 1, 2, 3


class WinGeom(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("B", "D", "H", "W", "wd", "wh", "ww", "sd", "sh", "sw",
                 "nwd", "nwh", "nww", "N", "Wd", "Wh", "Ww")]

    def rows(self) -> int:
        return self.B * self.nwd * self.nwh * self.nww * self.N

    def tokens(self) -> int:
        return self.B * self.D * self.H * self.W


class Epilogue(C.Structure):
    _fields_ = [
        ("cscale", C.c_void_p), ("bias", C.c_void_p),
        ("act", C.c_int32), ("ldm", C.c_int32),
        ("mul", C.c_void_p), ("resid", C.c_void_p),
        ("out_f32", C.c_void_p), ("out_bf16", C.c_void_p),
        ("ldo", C.c_int32), ("_pad", C.c_int32),
        ("win", C.POINTER(WinGeom)),
    ]


class LavtError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load the library (once). Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LavtError(
            f"{LIB_PATH} not found: build it with `python -m lavt_rs_b200.build` "
            "(the B200 path has no CPU / PyTorch fallback)")
    l = C.CDLL(LIB_PATH)
    l.lavt_last_error.restype = C.c_char_p
    _declare(l)
    if l.lavt_abi_version() != ABI_VERSION:
        raise LavtError("liblavt_b200.so ABI version mismatch; rebuild")
    _lib = l
    return l


ABI_VERSION = 1

_i32, _i64, _f32, _vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p
_EP = C.POINTER(Epilogue)
_WG = C.POINTER(WinGeom)

# name -> argtypes; every function returns int status.  Kept in the order of include/lavt_b200.h.
SIGNATURES = {
    "lavt_check_device": [],
    "lavt_gemm_bf16": [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _EP, _vp],
    "lavt_conv3x3_bf16": [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _i32, _EP, _vp],
}
EXPORTS = ["lavt_last_error", "lavt_abi_version", *SIGNATURES.keys()]


def _declare(l: C.CDLL) -> None:
    l.lavt_abi_version.restype = C.c_int
    for name, argtypes in SIGNATURES.items():
        fn = getattr(l, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().lavt_last_error().decode(errors="replace")
        raise LavtError(f"{what} failed (code {rc}): {msg}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise LavtError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise LavtError(f"{name} must be {dtype}, got {t.dtype}")
    if t.stride(-1) != 1:
        raise LavtError(f"{name} must be contiguous in its last dimension")


def make_epilogue(*, cscale=None, bias=None, act=ACT_NONE, mul=None, resid=None,
                  out_f32=None, out_bf16=None, ldo=None, win: Optional[WinGeom] = None) -> Epilogue:
    e = Epilogue()
    for name, t in (("cscale", cscale), ("bias", bias), ("resid", resid), ("out_f32", out_f32)):
        if t is not None:
            _req(t, torch.float32, name)
    for name, t in (("mul", mul), ("out_bf16", out_bf16)):
        if t is not None:
            _req(t, torch.bfloat16, name)
    e.cscale, e.bias = ptr(cscale), ptr(bias)
    e.act = act
    e.mul = ptr(mul)
    e.ldm = mul.stride(-2) if mul is not None else 0
    e.resid = ptr(resid)
    e.out_f32, e.out_bf16 = ptr(out_f32), ptr(out_bf16)
    out = out_f32 if out_f32 is not None else out_bf16
    if out is None:
        raise LavtError("epilogue needs out_f32 or out_bf16")
    e.ldo = int(ldo if ldo is not None else out.stride(-2))
    for t in (resid, out_f32, out_bf16):
        if t is not None and t.stride(-2) != e.ldo:
            raise LavtError("resid / out_f32 / out_bf16 must share one row pitch")
    e.win = C.pointer(win) if win is not None else None
    return e


def gemm_bf16(a: torch.Tensor, w: torch.Tensor, **epi) -> None:
    """out = epilogue(a @ w.T); a [M,K] bf16, w [N,K] bf16 (nn.Linear layout)."""
    _req(a, torch.bfloat16, "a")
    _req(w, torch.bfloat16, "w")
    M, K = a.shape
    N, K2 = w.shape
    if K != K2:
        raise LavtError(f"gemm: K mismatch {K} vs {K2}")
    e = make_epilogue(**epi)
    check(lib().lavt_gemm_bf16(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K,
                               C.byref(e), stream_ptr()), "lavt_gemm_bf16")


def conv3x3_bf16(x_nhwc: torch.Tensor, w_taps: torch.Tensor, **epi) -> None:
    """x [n_img,H,W,Cin] bf16 NHWC; w_taps [Cout, 9*Cin] bf16 (tap-major: (ky*3+kx)*Cin+ci)."""
    _req(x_nhwc, torch.bfloat16, "x")
    _req(w_taps, torch.bfloat16, "w_taps")
    n_img, H, W, Cin = x_nhwc.shape
    if not x_nhwc.is_contiguous():
        raise LavtError("conv: x must be contiguous NHWC")
    Cout = w_taps.shape[0]
    e = make_epilogue(**epi)
    check(lib().lavt_conv3x3_bf16(x_nhwc.data_ptr(), Cin, n_img, H, W, Cin, w_taps.data_ptr(), Cout,
                                  C.byref(e), stream_ptr()), "lavt_conv3x3_bf16")
