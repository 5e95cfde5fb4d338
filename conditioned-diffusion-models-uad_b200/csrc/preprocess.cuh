// GPU-resident volume preprocessing; see preprocess.cu for the reference call sites.
#pragma once
#include "common.h"

namespace cddpm {

// torchio.CropOrPad((h, w, d), padding_mode=pad) of a contiguous [H][W][D] volume (centre crop / pad; the odd voxel of
// a difference goes to the start).  crop_or_pad_offsets: source index of target index 0 along one axis.
void crop_or_pad_offsets(int source, int target, int* offset);
int launch_crop_or_pad(const float* in, int H, int W, int D, float* out, int h, int w, int d, float pad,
                       cudaStream_t stream);

// torchio.RescaleIntensity((out_min, out_max), percentiles=(p_lo, p_hi), masking_method=<mask>) in place on n voxels:
// np.percentile over the voxels with mask > 0, np.clip to the two cut-offs, then (x - min) / (max - min) * out_range +
// out_min with min / max of the clipped array.  An empty mask or a zero range leaves the volume unchanged.
// cutoffs (optional, device): the two float64 percentiles.
size_t rescale_workspace_bytes(int64_t n);
int launch_rescale_intensity(float* vol, const float* mask, int64_t n, double p_lo, double p_hi, float out_min,
                             float out_max, void* workspace, size_t workspace_bytes, double* cutoffs,
                             cudaStream_t stream);

// torchio.Resample(factor) of a unit-spacing [H][W][D] volume to spacing (fy, fx, fz): output extents
// ceil(N / f), sample positions 0.5 (f - 1) + f i.  bspline != 0: cubic B-spline (image_interpolation='bspline',
// float64 coefficients, mirror boundary); else nearest neighbour (label maps).  Samples outside the buffer read 0.
void resample_size(int source, double factor, int* target);
size_t resample_workspace_bytes(int H, int W, int D);
int launch_resample(const float* in, int H, int W, int D, double fy, double fx, double fz, int bspline, float* out,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace cddpm
