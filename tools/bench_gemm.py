"""Micro-benchmark of the tcgen05 GEMM / conv kernel at the hot-path shapes (CUDA events, L2-cold-ish:
operands are rotated through a ring larger than L2 when they are small)."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lavt_rs_b200 import _cabi  # noqa: E402


def timeit(fn, iters=20, warm=3):
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
    res = []
    shapes = [  # (M, N, K, tag)  B=1 clip 8x384^2 unless noted
        (73728, 384, 128, "s0 qkv"), (73728, 128, 128, "s0 proj"), (73728, 512, 128, "s0 fc1"), (73728, 128, 512, "s0 fc2"),
        (18432, 768, 256, "s1 qkv"), (18432, 1024, 256, "s1 fc1"), (18432, 256, 1024, "s1 fc2"),
        (4608, 1536, 512, "s2 qkv"), (4608, 512, 512, "s2 proj"), (4608, 2048, 512, "s2 fc1"), (4608, 512, 2048, "s2 fc2"),
        (1152, 3072, 1024, "s3 qkv"), (1152, 4096, 1024, "s3 fc1"), (1152, 1024, 4096, "s3 fc2"),
        (8192, 8192, 8192, "square 8k"), (36864, 2048, 512, "s2 fc1 B=8"),
    ]
    if os.environ.get("LAVT_BENCH_SHAPES") == "big":     # the 8-clip bench shapes of stage 2 (18 of the 24 blocks) + the square reference
        shapes = [(50176, 1536, 512, "s2 qkv B=8"), (50176, 512, 512, "s2 proj B=8"), (36864, 2048, 512, "s2 fc1 B=8"),
                  (36864, 512, 2048, "s2 fc2 B=8"), (147456, 1024, 256, "s1 fc1 B=8"), (8192, 8192, 8192, "square 8k")]
    for M, N, K, tag in shapes:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = torch.randn(N, K, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        bias = torch.zeros(N, device="cuda")
        t = timeit(lambda: _cabi.gemm_bf16(a, w, bias=bias, out_bf16=out))
        t_ref = timeit(lambda: torch.matmul(a, w.t()))
        res.append(dict(tag=tag, M=M, N=N, K=K, us=round(t * 1e6, 1), tflops=round(2 * M * N * K / t / 1e12, 1),
                        cublas_us=round(t_ref * 1e6, 1), cublas_tflops=round(2 * M * N * K / t_ref / 1e12, 1)))
        print(res[-1], flush=True)
        if os.environ.get("LAVT_BENCH_SHAPES") == "big" and K == 512 and N >= 1536:     # the same shape with the GELU epilogue (fc1)
            t = timeit(lambda: _cabi.gemm_bf16(a, w, bias=bias, act=_cabi.ACT_GELU, out_bf16=out))
            print(dict(tag=tag + " +GELU", us=round(t * 1e6, 1), tflops=round(2 * M * N * K / t / 1e12, 1)), flush=True)
    convs = [(8, 96, 640, "conv1_2"), (8, 96, 512, "conv2_2"), (8, 48, 768, "conv1_3"), (8, 24, 1536, "conv1_4")]
    if os.environ.get("LAVT_BENCH_SHAPES") == "big":
        convs = [(64, 96, 640, "conv1_2 B=8"), (64, 96, 512, "conv2_2 B=8"), (64, 48, 768, "conv1_3 B=8")]
    for n_img, H, Cin, tag in convs:
        x = torch.randn(n_img, H, H, Cin, device="cuda").bfloat16()
        w = torch.randn(512, 9 * Cin, device="cuda").bfloat16()
        out = torch.empty(n_img * H * H, 512, device="cuda", dtype=torch.bfloat16)
        sc = torch.ones(512, device="cuda")
        t = timeit(lambda: _cabi.conv3x3_bf16(x, w, cscale=sc, bias=sc, act=_cabi.ACT_RELU, out_bf16=out))
        fl = 2 * n_img * H * H * 512 * 9 * Cin
        xc = x.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        wc = w.view(512, 3, 3, Cin).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        t_ref = timeit(lambda: torch.nn.functional.conv2d(xc, wc, padding=1))
        res.append(dict(tag=tag, us=t * 1e6, tflops=fl / t / 1e12, cudnn_us=t_ref * 1e6, cudnn_tflops=fl / t_ref / 1e12))
        print(res[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_gemm.json", "w"), indent=1)


if __name__ == "__main__":
    main()
