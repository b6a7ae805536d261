import sys, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from oracle import lavt_oracle as O
from lavt_rs_b200 import _cabi as K, engine as E, train_engine as T
from lavt_rs_b200.lib.mask_predictor import SimpleDecoding
from lavt_rs_b200.weights import load_reference_state_dict
rel = lambda a, b: ((a.float().cpu() - b.float().cpu()).norm() / b.float().cpu().norm()).item()
cfg = O.OracleConfig(depths=(2, 2, 2, 2)); sd = O.random_state_dict(cfg, seed=0)
dec = SimpleDecoding(1024, None); load_reference_state_dict(dec, sd, "classifier."); dec = dec.cuda().train()
g = torch.Generator().manual_seed(21); n = 3
shapes = [(n, 1024, 3, 3), (n, 512, 6, 6), (n, 256, 12, 12), (n, 128, 24, 20)]
cs = [torch.randn(s, generator=g).to(torch.bfloat16).float() for s in shapes]
ws = E.workspace('cuda')
nhwc = [c.permute(0, 2, 3, 1).contiguous().cuda().to(torch.bfloat16) for c in cs]
lg, saved = T.decoder_fwd(dec, *nhwc, ws)
levels = saved[0]
q = O._q_bf16
y = cs[0]
for li, (skip, names) in enumerate(zip(cs[1:], (("conv1_4", "bn1_4", "conv2_4", "bn2_4"), ("conv1_3", "bn1_3", "conv2_3", "bn2_3"), ("conv1_2", "bn1_2", "conv2_2", "bn2_2")))):
    cat = torch.cat([O._up_to(y, skip), skip], 1)
    s1, s2 = levels[li][1], levels[li][2]
    print(li, 'cat', rel(s1[0].permute(0, 3, 1, 2), q(cat)))
    z1 = F.conv2d(q(cat), q(sd[f"classifier.{names[0]}.weight"]), padding=1)
    print(li, 'z1', rel(s1[1].view(n, cat.shape[2], cat.shape[3], -1).permute(0, 3, 1, 2), z1))
    t1 = O._cbr(cat, sd, names[0], names[1], True, True)
    print(li, 't1', rel(s1[3].permute(0, 3, 1, 2), q(t1)))
    mu = z1.mean((0, 2, 3)); rstd = 1 / torch.sqrt(z1.var((0, 2, 3), unbiased=False) + 1e-5)
    print(li, 'stats', rel(s1[2][0], mu), rel(s1[2][1], rstd))
    t2 = O._cbr(t1, sd, names[2], names[3], True, True)
    print(li, 't2', rel(s2[3].permute(0, 3, 1, 2), q(t2)))
    y = q(t2)
