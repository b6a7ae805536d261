import sys, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from lavt_rs_b200 import _cabi as K
g = torch.Generator().manual_seed(0)
for n, H, W, Cin, Cout in ((3, 12, 10, 512, 512), (2, 24, 24, 640, 512), (4, 7, 9, 1536, 512), (2, 96, 96, 640, 512), (1, 5, 3, 64, 128)):
    x = torch.randn(n, Cin, H, W, generator=g).to(torch.bfloat16)
    dz = torch.randn(n, Cout, H, W, generator=g).to(torch.bfloat16)
    w = torch.zeros(Cout, Cin, 3, 3, requires_grad=True)
    F.conv2d(x.float(), w, padding=1).backward(dz.float())
    ref = w.grad.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin)
    xs = x.permute(0, 2, 3, 1).contiguous().cuda(); dzs = dz.permute(0, 2, 3, 1).contiguous().cuda()
    dw = torch.full((Cout, 9 * Cin), 0.25, device='cuda')
    ws = torch.empty(K.conv3x3_wgrad_workspace_floats(n, H, W, Cin, Cout), device='cuda')
    try:
        K.conv3x3_wgrad(dzs, xs, dw, ws, accumulate=True)
        torch.cuda.synchronize()
    except Exception as e:
        print((n, H, W, Cin, Cout), 'FAILED', str(e)[:200]); break
    print((n, H, W, Cin, Cout), 'rel', ((dw.cpu() - 0.25 - ref).norm() / ref.norm()).item())
