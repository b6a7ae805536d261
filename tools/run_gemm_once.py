"""One GEMM shape a few times (ncu target): python tools/run_gemm_once.py M N K [act]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lavt_rs_b200 import _cabi as K_
M, N, K = (int(x) for x in sys.argv[1:4])
act = int(sys.argv[4]) if len(sys.argv) > 4 else 0
a = torch.randn(M, K, device="cuda").bfloat16()
w = torch.randn(N, K, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
bias = torch.zeros(N, device="cuda")
for _ in range(3):
    K_.gemm_bf16(a, w, bias=bias, act=act, out_bf16=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    K_.gemm_bf16(a, w, bias=bias, act=act, out_bf16=out)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 20 * 1e-3
print(f"M{M} N{N} K{K} act{act}: {t*1e6:.1f} us  {2.0*M*N*K/t/1e12:.1f} TF")
