import os, sys, torch
sys.path.insert(0, "/root/repo" if os.path.exists("/root/repo") else ".")
from lavt_rs_b200 import _cabi as K
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/it*1e3
for (n,H,ph,C1,C2) in [(64,96,48,512,128),(64,48,24,512,256),(64,24,12,1024,512)]:
    prev=torch.randn(n,ph,ph,C1,device="cuda").bfloat16(); skip=torch.randn(n,H,H,C2,device="cuda").bfloat16()
    out=torch.empty(n,H,H,C1+C2,device="cuda",dtype=torch.bfloat16)
    us=t(lambda: K.upsample_concat(prev,skip,out))
    byt=out.numel()*2+skip.numel()*2+prev.numel()*2
    print(f"upsample_concat {H}x{H} C{C1}+{C2}: {us:.1f} us  {byt/us/1e3:.0f} GB/s")
# reference points on the same buffers: a plain device copy of the output-sized tensor (read + write) and a fill (write only)
    src = torch.empty_like(out)
    us_c = t(lambda: out.copy_(src)); us_f = t(lambda: out.zero_())
    print(f"   copy of the output tensor: {us_c:.1f} us  {2*out.numel()*2/us_c/1e3:.0f} GB/s   fill: {us_f:.1f} us  {out.numel()*2/us_f/1e3:.0f} GB/s")
