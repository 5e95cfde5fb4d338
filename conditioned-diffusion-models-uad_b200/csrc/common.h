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
