// Host runtime shared by all kernels: last-error string, CUDA error mapping, TMA descriptor
// construction through the driver entry point (no link-time dependency on libcuda).
#include "common.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace lavt {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return LAVT_OK;
  set_last_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return LAVT_ERR_CUDA;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int make_tmap_typed(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz,
                           CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return LAVT_ERR_CUDA;
  }
  cuuint64_t gdims[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_last_error("TMA base pointer %p not 16-byte aligned", base);
    return LAVT_ERR_SHAPE;
  }
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr,
                  gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed: CUresult=%d (rank=%d dims0=%llu box0=%u)", (int)r, rank,
                   (unsigned long long)dims[0], box[0]);
    return LAVT_ERR_CUDA;
  }
  return LAVT_OK;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
  return make_tmap_typed(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swz);
}
// boxes whose inner extent is 64 bytes (one attention head of a [rows, 3C] tensor): no promotion, so the other half of the 128-byte
// line (the neighbouring head, consumed by another CTA much later) is not dragged through DRAM twice
int make_tmap_bf16_l2_64b(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box, CUtensorMapSwizzle swz) {
  return make_tmap_typed(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_64B);
}
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
  return make_tmap_typed(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, swz);
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    // measured on the 8-clip bench (331 launches per step replayed as one CUDA graph): 283.6 clips/s with the attribute on the GEMM /
    // LayerNorm / attention launches against 290.2 without -- the early-resident CTAs buy nothing next to persistent kernels that own
    // every SM until their last tile, so the attribute is opt-in (LAVT_PDL=1)
    const char* e = getenv("LAVT_PDL");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on != 0;
}

const char* last_error() { return g_err; }

}  // namespace lavt
