// extern "C" boundary (include/lavt_b200.h): argument validation + translation into kernel params.
#include "../../include/lavt_b200.h"
#include "common.cuh"
#include "gemm_tc.cuh"

#include <cstring>

namespace lavt {
const char* last_error();
int gemm_dispatch(const void* A, long long lda, const void* Bw, long long ldb, const GemmParams& p,
                  cudaStream_t stream);

static_assert(sizeof(lavt_win_geom_t) == sizeof(WinGeom), "public/private window geometry mismatch");

static int fill_epilogue(GemmParams& p, const lavt_epilogue_t* e) {
  LAVT_REQUIRE(e != nullptr, "epilogue pointer is NULL");
  p.cscale = e->cscale;
  p.bias = e->bias;
  p.act = e->act;
  p.mul = static_cast<const __nv_bfloat16*>(e->mul);
  p.ldm = e->ldm;
  p.resid = e->resid;
  p.out_f32 = e->out_f32;
  p.out_bf16 = static_cast<__nv_bfloat16*>(e->out_bf16);
  p.ldo = e->ldo;
  LAVT_REQUIRE(e->act >= 0 && e->act <= 3, "bad activation id %d", e->act);
  if (e->win) {
    p.rowmap = ROWMAP_WINDOW;
    std::memcpy(&p.win, e->win, sizeof(WinGeom));
  } else {
    p.rowmap = ROWMAP_IDENTITY;
  }
  return LAVT_OK;
}
}  // namespace lavt

using namespace lavt;

extern "C" {

const char* lavt_last_error(void) { return lavt::last_error(); }
int lavt_abi_version(void) { return LAVT_ABI_VERSION; }

int lavt_check_device(void) {
  int dev = 0;
  LAVT_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  LAVT_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_last_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
    return LAVT_ERR_ARCH;
  }
  return LAVT_OK;
}

int lavt_gemm_bf16(const void* A, int64_t lda, const void* Wt, int64_t ldw, int32_t M, int32_t N, int32_t K,
                   const lavt_epilogue_t* epi, void* stream) {
  GemmParams p;
  std::memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  int rc = fill_epilogue(p, epi);
  if (rc) return rc;
  if (p.rowmap == ROWMAP_WINDOW) {
    const WinGeom& g = p.win;
    long long rows = 1LL * g.B * g.nwd * g.nwh * g.nww * g.N;
    LAVT_REQUIRE(rows == M, "gemm: window geometry rows %lld != M %d", rows, M);
  }
  return gemm_dispatch(A, lda, Wt, ldw, p, static_cast<cudaStream_t>(stream));
}

int lavt_conv3x3_bf16(const void* x_nhwc, int64_t ldx, int32_t n_img, int32_t H, int32_t W, int32_t Cin,
                      const void* Wt, int32_t Cout, const lavt_epilogue_t* epi, void* stream) {
  GemmParams p;
  std::memset(&p, 0, sizeof(p));
  LAVT_REQUIRE(n_img > 0 && H > 0 && W > 0, "conv: empty input");
  p.M = n_img * H * W; p.N = Cout; p.K = 9 * Cin;
  int rc = fill_epilogue(p, epi);
  if (rc) return rc;
  LAVT_REQUIRE(p.rowmap == ROWMAP_IDENTITY, "conv: window row map not applicable");
  p.rowmap = ROWMAP_CONV;
  p.cH = H; p.cW = W; p.cCin = Cin; p.taps = 9;
  // 128-pixel tile: widest power-of-two strip that wastes the least padded area
  int best_tw = 8; long long best_area = -1;
  for (int tw = 8; tw <= 128; tw *= 2) {
    int th = 128 / tw;
    long long area = 1LL * ((H + th - 1) / th) * th * ((W + tw - 1) / tw) * tw;
    if (best_area < 0 || area < best_area || (area == best_area && tw > best_tw)) { best_area = area; best_tw = tw; }
  }
  p.cTW = best_tw; p.cTH = 128 / best_tw;
  p.cTilesW = (W + p.cTW - 1) / p.cTW;
  p.cTilesH = (H + p.cTH - 1) / p.cTH;
  return gemm_dispatch(x_nhwc, ldx, Wt, 9LL * Cin, p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
