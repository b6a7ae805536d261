import sys, torch
sys.path.insert(0, '.')
from lavt_rs_b200 import _cabi as K
g = torch.Generator().manual_seed(0)
M, N, Kk = 512, 512, 1296
slack = 64
a = torch.randn(M, Kk, generator=g).cuda().to(torch.bfloat16)
bfull = torch.randn(N, Kk, generator=g).cuda().to(torch.bfloat16)
for off in (0, 8, -8, 16, 1, -1, -19, 19):
    dst = torch.zeros(M, N, device='cuda')
    part = torch.empty(K.splitk_workspace_floats(M, N, Kk), device='cuda')
    try:
        K.gemm_bf16_splitk(a, bfull, dst, part, accumulate=False, b_koff=off)
        torch.cuda.synchronize()
    except Exception as e:
        print(off, 'FAILED', str(e)[:150]); break
    bs = torch.zeros(N, Kk, device='cuda')
    if off >= 0:
        bs[:, :Kk - off] = bfull[:, off:].float()
    else:
        bs[:, -off:] = bfull[:, :Kk + off].float()
    ref = a.float() @ bs.t()
    print(off, 'rel', ((dst - ref).norm() / ref.norm()).item())
