/* lavt_b200.h -- C ABI of the B200-native (sm_100a) LAVT-RS hot path.
 *
 * The reference (Yxxxb/LAVT-RS) has no FFI layer: its hot path is Python modules calling PyTorch ops.
 * Each entry point below replaces the PyTorch op sequence of one reference function (cited per
 * function as file:line relative to the reference root).  The Python host in lavt_rs_b200/lib/ binds
 * them with ctypes (lavt_rs_b200/_cabi.py) and keeps the reference's module / state-dict API.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; `stream` is a cudaStream_t passed as void*
 *   - activations: channels-last, tokens x C row-major; "bf16" = __nv_bfloat16, residual stream = fp32
 *   - weights: nn.Linear layout [out, in] (K-major), bf16
 *   - return value: 0 = ok, non-zero = error (LAVT_ERR_*); lavt_last_error() gives the message.
 *     There is NO CPU fallback: unsupported shapes / non-sm_100 devices are errors.
 */
#ifndef LAVT_B200_H_
#define LAVT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAVT_ABI_VERSION 1

#define LAVT_ERR_SHAPE 1
#define LAVT_ERR_CUDA 2
#define LAVT_ERR_ARCH 3

enum { LAVT_ACT_NONE = 0, LAVT_ACT_GELU = 1, LAVT_ACT_RELU = 2, LAVT_ACT_TANH = 3 };

/* Window geometry of one Swin block on a (B,D,H,W,C) channels-last tensor.
 * (wd,wh,ww)/(sd,sh,sw) are the EFFECTIVE window/shift after get_window_size clamping
 * (lib/video_swin_transformer.py:70-83); (Wd,Wh,Ww) is the configured window that sizes the
 * relative_position_bias_table (:107-127). nw* = ceil(dim / w*). N = wd*wh*ww. */
typedef struct lavt_win_geom {
  int32_t B, D, H, W;
  int32_t wd, wh, ww;
  int32_t sd, sh, sw;
  int32_t nwd, nwh, nww;
  int32_t N;
  int32_t Wd, Wh, Ww;
} lavt_win_geom_t;

/* Fused GEMM epilogue:
 *   out[orow(m), n] = act(acc[m,n] * cscale[n] + bias[n]) * mul[m,n] + resid[orow(m), n]
 * orow = m, or (win != NULL) the token row that window-row m maps back to (window_reverse +
 * reverse cyclic shift + crop, lib/video_swin_transformer.py:238-247); pad rows are dropped. */
typedef struct lavt_epilogue {
  const float* cscale;   /* [N] or NULL */
  const float* bias;     /* [N] or NULL */
  int32_t act;           /* LAVT_ACT_* */
  int32_t ldm;           /* row pitch of mul (elements) */
  const void* mul;       /* bf16 [M, ldm] or NULL */
  const float* resid;    /* fp32 [rows_out, ldo] or NULL */
  float* out_f32;        /* fp32 [rows_out, ldo] or NULL */
  void* out_bf16;        /* bf16 [rows_out, ldo] or NULL */
  int32_t ldo;           /* row pitch of resid / out (elements) */
  int32_t _pad;
  const lavt_win_geom_t* win; /* HOST pointer or NULL */
} lavt_epilogue_t;

const char* lavt_last_error(void);
int lavt_abi_version(void);
/* 0 if the current device is sm_100 (B200), LAVT_ERR_ARCH otherwise. */
int lavt_check_device(void);

/* C[M,N] = A[M,K] (bf16, pitch lda) x Wt[N,K]^T (bf16, pitch ldw), fp32 accumulate on tcgen05 tensor
 * cores, fused epilogue.  Replaces nn.Linear / Conv1d(k=1) call sites:
 *   qkv / proj  lib/video_swin_transformer.py:144,166 (lib/backbone.py:125,139)
 *   fc1 / fc2   :30-36          PatchMerging.reduction :309     PatchEmbed3D.proj :627
 *   PWAM vis_project / f_query / W / project_mm :900-973        LanguageGate res_gate :519-525
 * Requires N % 128 == 0, K % 64 == 0. */
int lavt_gemm_bf16(const void* A, int64_t lda, const void* Wt, int64_t ldw, int32_t M, int32_t N, int32_t K,
                   const lavt_epilogue_t* epi, void* stream);

/* 3x3 / pad 1 / stride 1 convolution as an implicit GEMM over NHWC bf16 input (pixel pitch ldx >= Cin):
 *   out[pix, co] = epilogue( sum_{ky,kx,ci} x[img, h+ky-1, w+kx-1, ci] * Wt[co, (ky*3+kx)*Cin + ci] )
 * Replaces conv{1,2}_{4,3,2} + BatchNorm2d(eval, folded into cscale/bias) + ReLU of
 * SimpleDecoding.forward, lib/mask_predictor.py:56-87.  Requires Cin % 64 == 0, Cout % 128 == 0. */
int lavt_conv3x3_bf16(const void* x_nhwc, int64_t ldx, int32_t n_img, int32_t H, int32_t W, int32_t Cin,
                      const void* Wt, int32_t Cout, const lavt_epilogue_t* epi, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAVT_B200_H_ */
