// GPU OpenSimplex-2D fractal noise (gen_noise, src/utils/generate_noise.py:8-52).
#pragma once
#include "common.h"

namespace cddpm {

// perm_host: 256-entry permutation (host memory).  out_f16: [B,1,H,W] half, the same field for every b (may be
// NULL); out_f32: [H,W] float (may be NULL).
int launch_simplex_noise(const unsigned char* perm_host, void* out_f16, float* out_f32, int B, int H, int W,
                         int octaves, double persistence, double frequency, cudaStream_t stream);

}  // namespace cddpm
