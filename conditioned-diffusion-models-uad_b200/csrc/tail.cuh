// Anomaly-scoring tail of the reference's _test_step (src/utils/utils_eval.py:18-194) on the GPU.
// Volumes are addressed logically as (y, x, d) = [H, W, D] like the reference's squeezed tensors; every input carries
// explicit element strides so both the dataloader layout [H,W,D] and the UNet output layout [D,1,H,W] are read in
// place.  Work buffers produced here are slice-major [D,H,W] (x fastest).
#pragma once
#include "common.h"

namespace cddpm {

struct VolView {
  const float* p = nullptr;
  int64_t sy = 0, sx = 0, sd = 0;
};

// sums[0..6] (double): sum|d|, sum d^2 over all voxels; over seg>0; over seg==0; [6] = count(seg>0)   (d = reco - orig)
// diff_masked[d][y][x] = eroded(mask)(y,x,d) ? |orig - reco| : 0, eroded = cross erosion x `iterations` per axial
// slice with zero border (apply_brainmask_volume, utils_eval.py:447-460); iterations < 1 = erode until stable (= all
// zero); erode == 0 skips the mask multiply.
int launch_residual_erode(const VolView& orig, const VolView& reco, const VolView& seg, const VolView& mask, int H,
                          int W, int D, int iterations, int erode, float* diff_masked, double* sums,
                          cudaStream_t stream);

// dst[Ho][Wo][Do] = trilinear resize of the [H,W,D] view with align_corners=True (F.interpolate, utils_eval.py:24-25).
// One slice of the output image grid of log_images (utils_eval.py:586-628): see compose_grid_kernel.
int launch_compose_grid(const float* panels, const float* ranges, int H, int W, uint8_t* rgb, cudaStream_t stream);
int launch_trilinear_resize(const VolView& src, int H, int W, int D, float* dst, int Ho, int Wo, int Do,
                            cudaStream_t stream);

// scipy.ndimage.median_filter(vol, (k,k,k), mode='reflect') for k = 5 (or 3) on a [D,H,W] buffer: symmetric padding,
// element of rank k^3/2 (utils_eval.py:462-464).
int launch_median3d(const float* in, float* out, int H, int W, int D, int k, cudaStream_t stream);

// max over the buffer (val_range top of find_best_val, utils_eval.py:86); out_max receives one float.
int launch_max(const float* x, int64_t n, float* out_max, cudaStream_t stream);

// counts[0] += #(g), counts[1+2i] += #(x > q[i]), counts[2+2i] += #(x > q[i] and g)  for i < nq (<= 4); g = seg > 0.
// x is a [D,H,W] work buffer; seg is read through its view.  Accumulating (caller zeroes) so that several volumes
// (and, across ranks, an all-reduce) can share one set of counters (global threshold, utils_eval.py:262-271).
int launch_threshold_counts(const float* x, const VolView& seg, int H, int W, int D, const float* q_host, int nq,
                            unsigned long long* counts, cudaStream_t stream);

// out[d][y][x] = x > thr (uint8)
int launch_threshold_mask(const float* x, int64_t n, float thr, unsigned char* out, cudaStream_t stream);

// Per image row y (the reference's "slice" loops run over axis 0, utils_eval.py:138-144, :160-174):
// rows[y][0] = #(x>thr), [1] = #g, [2] = #(x>thr and g), [3] = #(mask>0); rowsum[y] = sum of x over mask>0 (double).
int launch_row_stats(const float* x, const VolView& seg, const VolView& mask, int H, int W, int D, float thr,
                     unsigned long long* rows, double* rowsum, cudaStream_t stream);

// ROC-AUC and average precision of scores x (n values) against labels seg>0 with sklearn's tie semantics
// (compute_roc / compute_prc, utils_eval.py:548-557).  result[0] = AUC, result[1] = AP (double, device).
// workspace: ranking_workspace_bytes(n) bytes.
size_t ranking_workspace_bytes(int64_t n);
int launch_ranking_metrics(const float* x, const VolView& seg, int H, int W, int D, void* workspace,
                           size_t workspace_bytes, double* result, cudaStream_t stream);

// find_best_val (utils_eval.py:508-539) of the volume whose ranking pass last filled `ranking_workspace`: val_range =
// (0, max), `max_steps` quartile bisections decided on the device.  result[0] = best Dice, result[1] = its threshold, result[2] = max(x).
int launch_dice_bisect(const void* ranking_workspace, int64_t n, int max_steps, double* result, cudaStream_t stream);

// SURVEY.md §8 f-4 (tail_cc.cu): small-component filter, confusion counts, Hausdorff distance.  See include/cddpm_b200.h.
int launch_filter_small_components(const unsigned char* in, unsigned char* out, int H, int W, int D, int max_size,
                                   cudaStream_t stream);
int launch_confusion_counts(const unsigned char* pred, const VolView& seg, int H, int W, int D,
                            unsigned long long* counts, cudaStream_t stream);
size_t hausdorff_workspace_bytes(int H, int W, int D);
int launch_hausdorff(const unsigned char* pred, const VolView& seg, int H, int W, int D, void* workspace,
                     size_t workspace_bytes, long long* result, cudaStream_t stream);

}  // namespace cddpm
