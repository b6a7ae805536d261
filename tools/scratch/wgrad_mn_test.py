import sys, torch
sys.path.insert(0, '.')
from lavt_rs_b200 import _cabi as K
g = torch.Generator().manual_seed(0)
for M, N, Kin in ((1000, 96, 64), (4104, 384, 128), (392 * 9, 512, 2048), (3 * 147, 128, 128), (73728, 128, 512), (1152, 1024, 4096)):
    dy = torch.randn(M, N, generator=g).cuda().to(torch.bfloat16)
    x = torch.randn(M, Kin, generator=g).cuda().to(torch.bfloat16)
    dst = torch.full((N, Kin), 0.25, device='cuda')
    part = torch.empty(K.splitk_workspace_floats(N, Kin, M), device='cuda')
    try:
        K.gemm_bf16_wgrad(dy, x, dst, part, accumulate=True)
        torch.cuda.synchronize()
    except Exception as e:
        print(M, N, Kin, 'FAILED', str(e)[:200]); break
    ref = dy.float().t() @ x.float() + 0.25
    print(M, N, Kin, 'rel', ((dst - ref).norm() / ref.norm()).item())
