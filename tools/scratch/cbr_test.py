import sys, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from lavt_rs_b200 import _cabi as K, engine as E, train_engine as T
from lavt_rs_b200.lib.mask_predictor import SimpleDecoding
rel = lambda a, b: ((a.float().cpu() - b.float().cpu()).norm() / b.float().cpu().norm()).item()
g = torch.Generator().manual_seed(0)
dec = SimpleDecoding(1024, None).cuda().train()
ws = E.workspace('cuda')
n, H, W = 3, 12, 10
for conv_name, bn_name in (("conv2_3", "bn2_3"), ("conv1_3", "bn1_3")):
    conv, bn = getattr(dec, conv_name), getattr(dec, bn_name)
    Cin = conv.weight.shape[1]
    with torch.no_grad():
        conv.weight.copy_(conv.weight.to(torch.bfloat16).float())
        bn.weight.copy_(torch.rand(512, generator=g) + 0.5); bn.bias.copy_(torch.randn(512, generator=g) * 0.3)
    x = torch.randn(n, Cin, H, W, generator=g).to(torch.bfloat16).float()
    gout = torch.randn(n, 512, H, W, generator=g).to(torch.bfloat16).float()
    xr = x.clone().requires_grad_()
    wr = conv.weight.detach().cpu().clone().requires_grad_(); gr = bn.weight.detach().cpu().clone().requires_grad_(); br = bn.bias.detach().cpu().clone().requires_grad_()
    z = F.conv2d(xr, wr, padding=1)
    y = F.relu(F.batch_norm(z, None, None, gr, br, True, 0.1, 1e-5))
    y.backward(gout)
    grads = T.GradStore()
    xn = x.permute(0, 2, 3, 1).contiguous().cuda().to(torch.bfloat16)
    t, saved = T._cbr_fwd(xn, dec, conv_name, bn_name, ws, False)
    print(conv_name, 'fwd', rel(t.permute(0, 3, 1, 2), y))
    dx = torch.empty(n, H, W, Cin, device='cuda', dtype=torch.bfloat16)
    dt = gout.permute(0, 2, 3, 1).contiguous().cuda().to(torch.bfloat16).view(-1, 512)
    T._cbr_bwd(dec, saved, dt, grads, ws, False, dx)
    nm = grads.named(dec)
    print(' dx', rel(dx.permute(0, 3, 1, 2), xr.grad), 'dW', rel(nm[conv_name + '.weight'], wr.grad), 'dgamma', rel(nm[bn_name + '.weight'], gr.grad), 'dbeta', rel(nm[bn_name + '.bias'], br.grad))
    # pieces: dz
    zz = z.detach(); mu = zz.mean((0, 2, 3), keepdim=True); var = zz.var((0, 2, 3), unbiased=False, keepdim=True)
# upsample bwd
prev = torch.randn(n, 512, 6, 5, generator=g).requires_grad_()
up = F.interpolate(prev, size=(12, 10), mode='bilinear', align_corners=True)
go = torch.randn(n, 512 + 256, 12, 10, generator=g).to(torch.bfloat16).float()
up.backward(go[:, :512])
dprev = torch.empty(n, 6, 5, 512, device='cuda', dtype=torch.bfloat16)
K.upsample_concat_bwd(go.permute(0, 2, 3, 1).contiguous().cuda().to(torch.bfloat16), dprev)
print('upsample bwd', rel(dprev.permute(0, 3, 1, 2), prev.grad))
