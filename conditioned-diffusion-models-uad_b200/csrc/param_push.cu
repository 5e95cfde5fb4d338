// Whole-model parameter push as two launches of one table-driven kernel.
//
// After every optimizer step the engine re-lays every fp32 master parameter (the nn.Parameters of the reference's
// UNetModel, OpenAI_Unet.py:513-1006) into the layouts its kernels read: forward K-major 16-bit panels, transposed
// panels for the data gradient, plain fp32 copies (GroupNorm affine, biases) and a few bias sums.  One launch per
// parameter (~500 launches, even graph-replayed) costs ~1.7 ms of a 44 ms training step and the transposed pack reads
// with a stride of a whole filter per lane.  Here the per-parameter launches are RECORDED as jobs (common.h:
// ParamJob), cut into tiles, and executed by `param_push_kernel`: one CTA per tile, staged through shared memory so
// that both the fp32 reads and the 16-bit writes are contiguous runs.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <vector>

#include "common.h"

namespace cddpm {

struct ParamJobRecorder {
  std::vector<ParamJob> jobs;
};

static thread_local ParamJobRecorder* g_recorder = nullptr;

ParamJobRecorder* job_recorder() { return g_recorder; }
void job_record(const ParamJob& j) {
  if (g_recorder != nullptr) g_recorder->jobs.push_back(j);
}

int copy_f32_or_record(float* dst, const float* src, long long n, cudaStream_t stream) {
  if (g_recorder != nullptr) {
    ParamJob j = {kJobCopy, 0, 0, 0, 0, 0, 0, 0, 0, src, nullptr, dst, n};
    g_recorder->jobs.push_back(j);
    return kOk;
  }
  return check_cuda(cudaMemcpyAsync(dst, src, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDeviceToDevice, stream),
                    "param copy");
}

namespace {

constexpr int kTile = 32;          // 32 output channels x 32 input channels x taps per CTA
constexpr int kFlatChunk = 4096;   // elements per CTA for copy / add jobs
constexpr int kPushThreads = 256;
constexpr int kMaxTaps = 9;
constexpr int kRow = kTile * kMaxTaps + 1;  // +1: the transposed read walks rows with lane = co

struct PushTile {
  int job;
  int tile;
};

__device__ __forceinline__ uint16_t to16(float v, int fmt) {
  if (fmt == 1) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __half h = __float2half_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

__global__ void __launch_bounds__(kPushThreads) param_push_kernel(const ParamJob* __restrict__ jobs,
                                                                  const PushTile* __restrict__ tiles) {
  __shared__ float s[kTile * kRow];
  const PushTile t = tiles[blockIdx.x];
  const ParamJob j = jobs[t.job];
  const int tid = threadIdx.x;
  if (j.kind == kJobCopy || j.kind == kJobVecAdd) {
    const long long base = static_cast<long long>(t.tile) * kFlatChunk;
    const float* a = reinterpret_cast<const float*>(j.src);
    const float* b = reinterpret_cast<const float*>(j.src2);
    float* o = reinterpret_cast<float*>(j.dst);
#pragma unroll 4
    for (int k = tid; k < kFlatChunk; k += kPushThreads) {
      const long long i = base + k;
      if (i < j.n) o[i] = a[i] + (b != nullptr ? b[i] : 0.f);
    }
    return;
  }
  const int taps = j.ksize * j.ksize;
  const int ci_tiles = (j.c_s + kTile - 1) / kTile;
  const int co0 = (t.tile / ci_tiles) * kTile;
  const int ci0 = (t.tile % ci_tiles) * kTile;
  const int nco = min(kTile, j.cout - co0);
  const int nci = min(kTile, j.c_s - ci0);
  const int run = nci * taps;  // contiguous fp32 run per output channel: [ci0 .. ci0+nci) x taps
  const float* w = reinterpret_cast<const float*>(j.src);
  const int warp = tid >> 5, lane = tid & 31;
  for (int r = warp; r < nco; r += kPushThreads / 32) {
    const float* row = w + (static_cast<size_t>(co0 + r) * j.cin_total + j.cin_off + ci0) * taps;
    for (int k = lane; k < run; k += 32) s[r * kRow + k] = row[k];
  }
  __syncthreads();
  uint16_t* out = reinterpret_cast<uint16_t*>(j.dst);
  if (j.kind == kJobPack) {
    // out[co][koff + tap * C_s + ci]: lane = ci, one (co, tap) row per warp iteration
    for (int r = warp; r < nco * taps; r += kPushThreads / 32) {
      const int co = r / taps, tap = r % taps;
      if (lane < nci)
        out[static_cast<size_t>(co0 + co) * j.ktot + j.koff + tap * j.c_s + ci0 + lane] =
            to16(s[co * kRow + lane * taps + tap], j.fmt);
    }
  } else {
    // out_t[ci][koff + tap * Cout + co] = w[co][ci][taps - 1 - tap]: lane = co, one (ci, tap) row per warp iteration
    for (int r = warp; r < nci * taps; r += kPushThreads / 32) {
      const int ci = r / taps, tap = r % taps;
      if (lane < nco)
        out[static_cast<size_t>(ci0 + ci) * j.ktot + j.koff + tap * j.cout + co0 + lane] =
            to16(s[lane * kRow + ci * taps + (taps - 1 - tap)], j.fmt);
    }
  }
}

}  // namespace

struct ParamPushTable {
  ParamJob* d_jobs = nullptr;
  PushTile* d_tiles = nullptr;
  int n_first = 0, n_second = 0;  // tiles of phase 0 (packs, copies) and of phase 1 (sums over copied vectors)
};

void param_push_table_free(ParamPushTable* t) {
  if (t == nullptr) return;
  if (t->d_jobs != nullptr) cudaFree(t->d_jobs);
  if (t->d_tiles != nullptr) cudaFree(t->d_tiles);
  delete t;
}

int param_push_record_begin(ParamJobRecorder** rec) {
  if (g_recorder != nullptr) return fail(kInvalidArgument, "param push: recorder already installed on this thread");
  *rec = new ParamJobRecorder();
  g_recorder = *rec;
  return kOk;
}

int param_push_record_end(ParamJobRecorder* rec, ParamPushTable** out, bool ok) {
  g_recorder = nullptr;
  std::vector<ParamJob> jobs;
  jobs.swap(rec->jobs);
  delete rec;
  *out = nullptr;
  if (!ok) return kOk;
  std::vector<PushTile> first, second;
  for (size_t i = 0; i < jobs.size(); ++i) {
    const ParamJob& j = jobs[i];
    int tiles;
    if (j.kind == kJobCopy || j.kind == kJobVecAdd) {
      tiles = static_cast<int>((j.n + kFlatChunk - 1) / kFlatChunk);
    } else {
      if (j.ksize * j.ksize > kMaxTaps) return fail(kInvalidArgument, "param push: kernel size above 3x3");
      tiles = ((j.cout + kTile - 1) / kTile) * ((j.c_s + kTile - 1) / kTile);
    }
    std::vector<PushTile>& dst = j.kind == kJobVecAdd ? second : first;
    for (int t = 0; t < tiles; ++t) dst.push_back(PushTile{static_cast<int>(i), t});
  }
  ParamPushTable* tab = new ParamPushTable();
  tab->n_first = static_cast<int>(first.size());
  tab->n_second = static_cast<int>(second.size());
  first.insert(first.end(), second.begin(), second.end());
  int st = kOk;
  if (!jobs.empty()) {
    st = check_cuda(cudaMalloc(&tab->d_jobs, jobs.size() * sizeof(ParamJob)), "param push: cudaMalloc");
    if (st == kOk) st = check_cuda(cudaMalloc(&tab->d_tiles, first.size() * sizeof(PushTile)), "param push: cudaMalloc");
    if (st == kOk)
      st = check_cuda(cudaMemcpy(tab->d_jobs, jobs.data(), jobs.size() * sizeof(ParamJob), cudaMemcpyHostToDevice),
                      "param push: job table upload");
    if (st == kOk)
      st = check_cuda(cudaMemcpy(tab->d_tiles, first.data(), first.size() * sizeof(PushTile), cudaMemcpyHostToDevice),
                      "param push: tile table upload");
  }
  if (st != kOk) {
    param_push_table_free(tab);
    return st;
  }
  *out = tab;
  return kOk;
}

int param_push_launch(const ParamPushTable* t, cudaStream_t stream) {
  if (t->n_first > 0) {
    param_push_kernel<<<t->n_first, kPushThreads, 0, stream>>>(t->d_jobs, t->d_tiles);
    CDDPM_TRY(check_launch("param_push_kernel"));
  }
  if (t->n_second > 0) {
    param_push_kernel<<<t->n_second, kPushThreads, 0, stream>>>(t->d_jobs, t->d_tiles + t->n_first);
    CDDPM_TRY(check_launch("param_push_kernel (sums)"));
  }
  return kOk;
}

}  // namespace cddpm
