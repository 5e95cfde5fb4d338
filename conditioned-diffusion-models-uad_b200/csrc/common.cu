#include "common.h"

#include <cudaTypedefs.h>

#include <stdlib.h>

#include <mutex>

namespace cddpm {

static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* last_error_cstr() { return g_last_error.c_str(); }

int fail(Status s, const std::string& msg) {
  set_last_error(msg);
  return s;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return kOk;
  return fail(kCudaError, std::string(what) + ": " + cudaGetErrorName(e) + ": " + cudaGetErrorString(e));
}

int check_launch(const char* what) { return check_cuda(cudaGetLastError(), what); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

int encode_tmap_16bit(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(kCudaError, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdims[5];
  cuuint64_t gstrides[5];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstrides[i] = strides_bytes[i];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                  gdims, gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(kCudaError, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
  }
  return kOk;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CDDPM_PDL");  // A/B switch for measurements: 0 = plain stream-ordered launches
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
static thread_local bool g_pdl_next = false;
void pdl_set_next(bool on) { g_pdl_next = on && pdl_enabled(); }
// Which kernels accept the programmatic launch that capture offers them.  Measured on one box (profiles/r02_summary.md,
// call 52-54): only the convolutions (default, 3) - their barrier init, TMEM allocation, descriptor prefetch and weight
// ring then run under the tail of the GroupNorm pass in front of them: B=32 forward 4.956 -> 4.784 ms, B=8 1.733 ->
// 1.661, B=1 1.187 -> 1.156; letting the GroupNorm / attention / stem / head kernels start early as well (1) gives the
// gain back at every batch size above 4 (their early blocks sit on SM slots the convolution's last tiles still use).
int pdl_mode() {
  static const int mode = [] {
    const char* e = getenv("CDDPM_PDL_MODE");  // 1 = every kernel with a griddepcontrol.wait, 2 = all but the convolutions
    return e != nullptr ? atoi(e) : 3;
  }();
  return mode;
}
bool pdl_elementwise_ok() { return pdl_mode() != 3; }
bool pdl_take_next() {
  const bool v = g_pdl_next;
  g_pdl_next = false;
  return v;
}

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace cddpm
