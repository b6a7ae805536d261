// Parameter block of the tcgen05 GEMM / implicit-GEMM-conv kernel (gemm_tc.cu).
#pragma once
#include "geom.cuh"
#include <cuda_bf16.h>

namespace lavt {

enum GemmAct { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_TANH = 3, ACT_SIGMOID = 4 };
enum GemmRowMap { ROWMAP_IDENTITY = 0, ROWMAP_WINDOW = 1, ROWMAP_CONV = 2, ROWMAP_WGCONV = 3 };

// pre = acc[m,n] * cscale[n] + bias[n];  out_pre[orow(m), n] = pre
// out[orow(m), n] = act(pre) * mul'[m,n] * rscale[orow(m) / rs_rows] + resid[orow(m), n],  mul' = mul or GELU'(mul) (mul_act)
struct GemmParams {
  int M, N, K;               // logical problem (conv: M = pixels, K = taps*Cin)
  // ---- epilogue ----
  const float* cscale;       // [N] or nullptr
  const float* bias;         // [N] or nullptr
  int act;                   // GemmAct
  const __nv_bfloat16* mul;  // [M, ldm] or nullptr (indexed by the GEMM row m)
  int ldm;
  int mul_act;               // 0: multiply by mul; ACT_GELU: multiply by GELU'(mul) (fc2 input gradient x derivative of the fc1 pre-activation)
  __nv_bfloat16* out_pre;    // [rows_out, ldo] or nullptr: the value before the activation (training forward)
  int pre_mode;              // 1 (with act = ACT_GELU): out_pre receives GELU'(pre) instead of pre -- what the backward of fc1 multiplies by
  const float* resid;        // [rows_out, ldo] or nullptr (indexed by the OUTPUT row)
  float* out_f32;            // [rows_out, ldo] or nullptr
  __nv_bfloat16* out_bf16;   // [rows_out, ldo] or nullptr
  int ldo;
  const float* rscale;       // per-sample scale of the branch (stochastic depth in training), indexed by output row / rs_rows; or nullptr
  int rs_rows;
  // ---- row map ----
  int rowmap;                // GemmRowMap
  WinGeom win;               // ROWMAP_WINDOW
  // ROWMAP_CONV: A is a 4-D NHWC tensor (C, W, H, Nimg), 3x3 taps, pad 1
  int cH, cW, cCin;          // image height/width, input channels
  int cTH, cTW;              // tile = cTH x cTW pixels (cTH*cTW == 128)
  int cTilesH, cTilesW;      // tiles per image
  int taps;                  // 9 (3x3), 27 (3x3x3) or 1
  int cD;                    // frames per clip for the 3-D conv (A is (C, W, H, D, Nclip)); 0 / 1 = 2-D
  // ---- split-K (weight gradients: few output tiles, very long K) ----
  int ksplit;                // 0 / 1 = off; s > 1: work item = (tile, split); split ks accumulates k-blocks [ks*kbs, (ks+1)*kbs)
  int kbs;                   // k-blocks per split
  long long split_stride;    // out_f32 of split ks = out_f32 + ks * split_stride (partials, reduced by splitk_reduce)
  // ROWMAP_WGCONV (conv3x3 WEIGHT gradient, with mnmajor = 1): C[co, tap*Cin + ci] = sum_pixels dz[pix, co] * x[pix + tap, ci].  A = dz
  // and B = x are 4-D NHWC tensor maps (C, W, H, Nimg); a k-block is a cTH x cTW = 64-pixel tile, the tap is a coordinate offset of
  // the B box and the zero padding is TMA's out-of-bounds fill -- no im2col, no padded or transposed copies.  M = Cout, N = 9*Cin.
  int mnmajor;               // 1: both operands are stored [K, M] / [K, N] row-major (dW = dY^T X straight from the row-major activations):
                             //    TMA boxes of 64 k-rows x 64 MN-elements, MN-major shared-memory descriptors, no transposed copies
  int b_koff;                // added to the K coordinate of the B operand (conv weight gradient: tap offset in the padded pixel axis)
};
int splitk_epilogue_dispatch(const float* partials, int splits, long long M, int N, const float* cscale, const float* bias, int act,
                             const float* resid, float* out_f32, __nv_bfloat16* out_bf16, long long ldo, cudaStream_t st);
int splitk_reduce_dispatch(const float* partials, int splits, long long count, int ncols, float* dst, long long ldd, int accumulate,
                           cudaStream_t st);

}  // namespace lavt
