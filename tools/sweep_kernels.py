"""BASELINE.json configs[4] -- kernel sweep: window attention + PWAM at stages 0-3, T in {4,8,16}, Nl in {20,22,40,77},
windows (8,7,7) and (8,12,12), Swin-B 384x384, 1 clip.  Reports CUDA-event time, achieved TFLOP/s, fraction of the measured
bf16 peak, and parity of each launch against a plain fp32 PyTorch evaluation of the same op.
    python tools/sweep_kernels.py  ->  gpurun_out/kernel_sweep.json"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lavt_rs_b200 import _cabi as K  # noqa: E402
from lavt_rs_b200 import engine as E  # noqa: E402
from lavt_rs_b200.geometry import window_geometry  # noqa: E402
from test_attention_gpu import torch_window_attention  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


def main():
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1590.0
    out = {"attention": [], "pwam": [], "peak_tflops_burst": peak}
    for window in ((8, 7, 7), (8, 12, 12)):
        for T in (4, 8, 16):
            for s in range(4):
                C, nH, HW = 128 * 2 ** s, 4 * 2 ** s, 96 // 2 ** s
                geom = window_geometry(1, T, HW, HW, window, True, True)
                rows = geom.rows()
                qkv = torch.randn(rows, 3 * C, device="cuda")
                qkv[:, :C] *= 32 ** -0.5 * math.log2(math.e)
                qkv = qkv.bfloat16()
                L = (2 * window[0] - 1) * (2 * window[1] - 1) * (2 * window[2] - 1)
                table = torch.randn(L, nH, device="cuda") * 0.5
                tt = table.t().contiguous()
                o = torch.empty(rows, C, device="cuda", dtype=torch.bfloat16)
                t = timeit(lambda: K.window_attention(qkv, tt, geom, o))
                ref = torch_window_attention(qkv, table, geom)
                rel = ((o.float() - ref).norm() / ref.norm()).item()
                fl = 4.0 * rows * geom.N * C
                out["attention"].append(dict(window=window, T=T, stage=s, N=geom.N, rows=rows, us=t * 1e6, tflops=fl / t / 1e12,
                                             frac_of_peak=fl / t / 1e12 / peak, rel_l2=rel))
                print(out["attention"][-1], flush=True)
                del ref
    # PWAM + gate chain (engine.pwam_gate) per stage
    from lavt_rs_b200.lib.video_swin_transformer import MMBasicLayer, PatchMerging
    from lavt_rs_b200.args import default_args
    sys.path.insert(0, ROOT)
    from oracle import lavt_oracle as O
    for s in range(4):
        C = 128 * 2 ** s
        layer = MMBasicLayer(dim=C, depth=2, num_heads=4 * 2 ** s, window_size=(8, 7, 7), qkv_bias=True, args=None).cuda().eval()
        torch.nn.init.trunc_normal_(layer.res_gate[0].weight, std=0.02)
        torch.nn.init.trunc_normal_(layer.res_gate[2].weight, std=0.02)
        sd = {f"backbone.layers.{s}." + k: v.detach().float().cpu() for k, v in layer.state_dict().items()}
        for T in (4, 8, 16):
            n = T * (96 // 2 ** s) ** 2
            for Nl in (20, 22, 40, 77):
                g = torch.Generator().manual_seed(s * 100 + T + Nl)
                x = torch.randn(1, n, C, generator=g)
                l = torch.randn(1, 768, Nl, generator=g)
                m = torch.zeros(1, Nl, 1)
                m[:, : math.ceil(0.7 * Nl)] = 1
                xd, ld, md = x.cuda().view(n, C).contiguous(), l.cuda(), m.cuda().view(1, Nl)
                xb = xd.to(torch.bfloat16)
                r = torch.empty(n, C, device="cuda")
                ws = E.workspace("cuda")
                x_work = xd.clone()

                def run():
                    x_work.copy_(xd)
                    E.pwam_gate(x_work, xb, layer.fusion, layer.res_gate, ld, md, 1, ws, r_f32=r)
                t = timeit(run)
                ref_r = O.pwam(x, l, m, sd, f"backbone.layers.{s}.fusion.", 1)
                ref_x = O.language_gate(x, ref_r, sd, f"backbone.layers.{s}.res_gate.")
                rel_r = ((r.cpu() - ref_r[0]).norm() / ref_r.norm()).item()
                rel_x = ((x_work.cpu() - ref_x[0]).norm() / ref_x.norm()).item()
                fl = 12.0 * n * C * C + 4.0 * Nl * 768 * C + 4.0 * n * Nl * C
                out["pwam"].append(dict(stage=s, T=T, Nl=Nl, n=n, us=t * 1e6, tflops=fl / t / 1e12, rel_l2_residual=rel_r, rel_l2_gated=rel_x))
                print(out["pwam"][-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/kernel_sweep.json", "w"), indent=1)


if __name__ == "__main__":
    main()
