// Host-side helpers shared by all translation units of libcddpm_b200: error reporting, launch checks, tensor-map
// encoding.  Nothing here is exported; the C-ABI lives in include/cddpm_b200.h and api.cu.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace cddpm {

typedef __nv_bfloat16 bf16;

// Status codes returned by every extern "C" entry point.
enum Status : int {
  kOk = 0,
  kInvalidArgument = 1,
  kCudaError = 2,
  kUnsupported = 3,
  kNotReady = 4,
};

void set_last_error(const std::string& msg);
int fail(Status s, const std::string& msg);
int check_cuda(cudaError_t e, const char* what);
int check_launch(const char* what);

#define CDDPM_CUDA(expr)                                  \
  do {                                                    \
    int _st = ::cddpm::check_cuda((expr), #expr);         \
    if (_st != ::cddpm::kOk) return _st;                  \
  } while (0)

#define CDDPM_TRY(expr)                  \
  do {                                   \
    int _st = (expr);                    \
    if (_st != ::cddpm::kOk) return _st; \
  } while (0)

// Encode a tiled tensor map (bf16/f16 elements, 128-byte swizzle, zero fill out of bounds).
// dims/box are innermost-first; strides_bytes has rank-1 entries (dims 1..rank-1).
int encode_tmap_16bit(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128 = true);

int device_sm_count();

// ---- Programmatic dependent launch (PDL).  A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization
// may start while its stream predecessor is still draining: its CTAs are placed as the predecessor's CTAs retire, run
// their prologue (barrier init, TMEM allocation, tensor-map prefetch, weight loads) and block in
// `griddepcontrol.wait` until the predecessor has completed and flushed.  The engines set the thread-local flag in
// front of an op whose stream predecessor is a kernel; launch_k consumes it.  ONLY kernels that execute pdl_wait()
// before their first dependent access may be launched through launch_k.  CDDPM_PDL=0 disables the attribute.
bool pdl_enabled();
void pdl_set_next(bool on);
bool pdl_take_next();
int pdl_mode();             // CDDPM_PDL_MODE: 3 (default) only the convolutions take the offer, 1 every kernel, 2 all but them
bool pdl_elementwise_ok();

template <typename... P, typename... A>
inline cudaError_t launch_k(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  int n = 0;
  if (pdl_take_next() && pdl_elementwise_ok()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

#ifdef __CUDACC__
// Device side of PDL: no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// Parameter re-layout jobs.  While a recorder is installed (thread-local), the small re-layout launches behind
// cddpm_unet_set_param (weight packing, plain copies, bias sums) append a job here instead of launching, so that a
// whole-model push becomes two launches of one table-driven kernel (param_push.cu).
enum ParamJobKind : int { kJobPack = 0, kJobPackT = 1, kJobCopy = 2, kJobVecAdd = 3 };
struct ParamJob {
  int kind;
  int cout, cin_total, ksize, cin_off, c_s, ktot, koff, fmt;
  const void* src;
  const void* src2;
  void* dst;
  long long n;  // elements
};
struct ParamJobRecorder;
ParamJobRecorder* job_recorder();              // nullptr when launches run directly
void job_record(const ParamJob& j);
// fp32 device copy that honours the recorder
int copy_f32_or_record(float* dst, const float* src, long long n, cudaStream_t stream);
// Record (begin .. end) the jobs of a whole-model push into a device-resident tile table, then replay it with two
// launches per push.  `ok == false` at end discards the recording.
struct ParamPushTable;
int param_push_record_begin(ParamJobRecorder** rec);
int param_push_record_end(ParamJobRecorder* rec, ParamPushTable** out, bool ok);
int param_push_launch(const ParamPushTable* t, cudaStream_t stream);
void param_push_table_free(ParamPushTable* t);

}  // namespace cddpm
