// Length regulator (first half): expand per-token features by the integer durations the
// predictor produced — frames[b, f, :] = feats[b, tok(b, f), :] with tok = the token whose cumulative-duration
// interval contains frame f — entirely on device: the prefix sum of the durations never leaves the GPU.
//
// One CTA = LR_FRAMES consecutive frames of one utterance.  Every CTA first rebuilds the utterance's inclusive
// cumulative durations in shared memory (T <= 1024 ints: a block scan is cheaper than a second kernel and a
// round trip through HBM), then each warp binary-searches its frame's token and copies the row with 128-bit
// accesses.  Frames past the utterance's total length are written as zeros (the output needs no prior memset).
// HBM-bound integer / copy work: algorithmic bytes = 4 C per frame written (+ the feature rows, read once from L2).
#pragma once
#include "elementwise.cuh"

namespace stz {

constexpr int LR_FRAMES = 32, LR_THREADS = 256, LR_MAX_T = 1024;

__global__ void __launch_bounds__(LR_THREADS) length_regulate_kernel(const float* __restrict__ feats, const int32_t* __restrict__ dur,
                                                                      float* __restrict__ frames, int32_t* __restrict__ frame_lens,
                                                                      int32_t* __restrict__ frame_tok, int T, int C, int F_max) {
  __shared__ int cum[LR_MAX_T];
  __shared__ int warp_tot[LR_THREADS / 32];
  pdl_sync();
  const int b = blockIdx.y, f0 = blockIdx.x * LR_FRAMES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // inclusive scan of dur[b, :] (negative entries count as 0): thread -> 4 consecutive tokens
  const int per = (T + LR_THREADS - 1) / LR_THREADS;      // <= 4
  int local[4], run = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = tid * per + i;
    int v = (i < per && t < T) ? dur[static_cast<size_t>(b) * T + t] : 0;
    run += v > 0 ? v : 0;
    local[i] = run;
  }
  int incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  int base = incl - run;
  for (int w = 0; w < warp; ++w) base += warp_tot[w];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = tid * per + i;
    if (i < per && t < T) cum[t] = base + local[i];
  }
  __syncthreads();
  const int total = T > 0 ? cum[T - 1] : 0;
  const int n_frames = total < F_max ? total : F_max;
  if (blockIdx.x == 0 && tid == 0) frame_lens[b] = n_frames;
  const int c4 = C >> 2;
  for (int fi = warp; fi < LR_FRAMES; fi += LR_THREADS / 32) {
    const int f = f0 + fi;
    if (f >= F_max) break;
    float4* dst = reinterpret_cast<float4*>(frames + (static_cast<size_t>(b) * F_max + f) * C);
    if (f >= n_frames) {
      for (int c = lane; c < c4; c += 32) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (frame_tok != nullptr && lane == 0) frame_tok[static_cast<size_t>(b) * F_max + f] = -1;
      continue;
    }
    int lo = 0, hi = T - 1;                      // first token with cum[t] > f
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cum[mid] > f) hi = mid; else lo = mid + 1;
    }
    const float4* src = reinterpret_cast<const float4*>(feats + (static_cast<size_t>(b) * T + lo) * C);
    for (int c = lane; c < c4; c += 32) dst[c] = __ldg(src + c);
    if (frame_tok != nullptr && lane == 0) frame_tok[static_cast<size_t>(b) * F_max + f] = lo;
  }
}

}  // namespace stz
