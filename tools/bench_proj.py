"""The proj GEMM of a stage-2 Swin block (M 50176, N 512, K 512) under its epilogue variants: plain bf16 out, fp32 out, + residual, + window-reverse
scatter (the real one).  python tools/bench_proj.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lavt_rs_b200 import _cabi as K
from lavt_rs_b200.geometry import window_geometry
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3
geom = window_geometry(8, 8, 24, 24, (8, 7, 7), True, True)
M, N, Kd = geom.rows(), 512, 512
tokens = geom.tokens()
a = torch.randn(M, Kd, device="cuda").bfloat16(); w = torch.randn(N, Kd, device="cuda").bfloat16(); bias = torch.zeros(N, device="cuda")
ob = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); of = torch.empty(M, N, device="cuda"); x = torch.randn(tokens, N, device="cuda")
xm = torch.randn(M, N, device="cuda")
flops = 2.0 * M * N * Kd
for name, fn in (("bf16 out", lambda: K.gemm_bf16(a, w, bias=bias, out_bf16=ob)), ("fp32 out", lambda: K.gemm_bf16(a, w, bias=bias, out_f32=of)),
                 ("fp32 out + residual (rows in place)", lambda: K.gemm_bf16(a, w, bias=bias, resid=xm, out_f32=xm)),
                 ("fp32 out + residual + window-reverse scatter", lambda: K.gemm_bf16(a, w, bias=bias, resid=x, out_f32=x, win=geom))):
    us = t(fn)
    print(f"{name:48s} {us:7.1f} us  {flops/us/1e6:7.1f} TF")
