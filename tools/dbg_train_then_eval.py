"""Does an eval forward change after training steps (state that should not leak)?  python tools/dbg_train_then_eval.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
model = B.build_model(False, dev)
x, ids, m = B.synth_batch(1, 5)
x, ids, m = x.to(dev), ids.to(dev), m.to(dev)
def feats():
    with torch.no_grad():
        l = model.text_encoder(ids, attention_mask=m)[0] if not hasattr(model, "encode_text") else None
        out = model(x, ids, m).float()
    return out
snap = {k: v.clone() for k, v in model.state_dict().items()}
y0 = feats()
B.train_step_extra(model, dev, 1, 0, None, clips=2, steps=1, warmup=1)
changed = [k for k, v in model.state_dict().items() if not torch.equal(v, snap[k])]
print("state entries changed by the training steps:", len(changed), changed[:6])
y1 = feats()
print("eval output after training (BN statistics moved): rel diff", ((y1 - y0).norm() / y0.norm()).item())
with torch.no_grad():
    for k, v in model.state_dict().items():
        if not torch.equal(v, snap[k]):
            v.copy_(snap[k])
y2 = feats()
print("eval output after restoring the state dict: rel diff", ((y2 - y0).norm() / y0.norm()).item())
from lavt_rs_b200 import engine as E
for mod in model.modules():
    if hasattr(mod, "prepared"):
        try: mod.prepared.clear()
        except Exception: pass
y3 = feats()
print("... and after clearing every prepared-weight cache: rel diff", ((y3 - y0).norm() / y0.norm()).item())
