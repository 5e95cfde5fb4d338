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

}  // namespace cddpm
