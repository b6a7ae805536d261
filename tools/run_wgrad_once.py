"""One weight-gradient GEMM (MN-major operands, split-K) and one input-gradient GEMM at the stage-2 fc1 shape of the training bench
(4 clips: 18 432 tokens, 512 -> 2048) for an ncu capture:  ncu --set full -k regex:gemm_bf16_tc python tools/run_wgrad_once.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lavt_rs_b200 import _cabi as K  # noqa: E402

tokens, n_in, n_out = 18432, 512, 2048
g = torch.Generator().manual_seed(0)
dy = torch.randn(tokens, n_out, generator=g).cuda().to(torch.bfloat16)
x = torch.randn(tokens, n_in, generator=g).cuda().to(torch.bfloat16)
w_t = torch.randn(n_in, n_out, generator=g).cuda().to(torch.bfloat16)       # transposed weight for dX = dY W
dw = torch.zeros(n_out, n_in, device="cuda")
part = torch.empty(K.splitk_workspace_floats(n_out, n_in, tokens), device="cuda")
dx = torch.empty(tokens, n_in, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    K.gemm_bf16_wgrad(dy, x, dw, part, accumulate=True)
    K.gemm_bf16(dy, w_t, out_bf16=dx)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn, fl in (("wgrad", lambda: K.gemm_bf16_wgrad(dy, x, dw, part, accumulate=True), 2.0 * tokens * n_in * n_out),
                     ("dgrad", lambda: K.gemm_bf16(dy, w_t, out_bf16=dx), 2.0 * tokens * n_in * n_out)):
    s.record()
    for _ in range(20):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = s.elapsed_time(e) / 20 * 1e-3
    print(f"{name}: {t*1e6:.1f} us, {fl / t / 1e12:.0f} TFLOP/s")
