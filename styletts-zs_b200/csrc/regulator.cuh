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

// feats2 != nullptr: a frame row is [feats row (C columns) | feats2 row (C2 columns)] — the prosody heads regulate the
// duration encoder's output and the per-token style summary in one pass, without a concatenated copy.
__global__ void __launch_bounds__(LR_THREADS) length_regulate_kernel(const float* __restrict__ feats, const int32_t* __restrict__ dur,
                                                                      float* __restrict__ frames, int32_t* __restrict__ frame_lens,
                                                                      int32_t* __restrict__ frame_tok, int T, int C, int F_max,
                                                                      const float* __restrict__ feats2, int C2) {
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
  const int c4 = C >> 2, c24 = feats2 != nullptr ? C2 >> 2 : 0;
  for (int fi = warp; fi < LR_FRAMES; fi += LR_THREADS / 32) {
    const int f = f0 + fi;
    if (f >= F_max) break;
    float4* dst = reinterpret_cast<float4*>(frames + (static_cast<size_t>(b) * F_max + f) * (C + 4 * c24));
    if (f >= n_frames) {
      for (int c = lane; c < c4 + c24; c += 32) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
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
    if (c24 > 0) {
      const float4* src2 = reinterpret_cast<const float4*>(feats2 + (static_cast<size_t>(b) * T + lo) * C2);
      for (int c = lane; c < c24; c += 32) dst[c4 + c] = __ldg(src2 + c);
    }
    if (frame_tok != nullptr && lane == 0) frame_tok[static_cast<size_t>(b) * F_max + f] = lo;
  }
}

// ---- prosody (F0 / energy) heads behind the regulator (second half) ---------------------------

// perm = sequence indices sorted by frame count, longest first (stable): the schedule of the BiLSTM recurrence kernel
__global__ void __launch_bounds__(1024) perm_from_lens_kernel(const int* __restrict__ lens, int* __restrict__ perm, int B) {
  pdl_sync();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int lb = lens[b];
    int rank = 0;
    for (int j = 0; j < B; ++j) {
      const int lj = lens[j];
      rank += (lj > lb || (lj == lb && j < b)) ? 1 : 0;
    }
    perm[rank] = b;
  }
}

// needed[t] = 1 iff the 128-row tile t of the [B * F] frame rows holds a frame below its utterance's frame count
__global__ void __launch_bounds__(256) frame_tile_needed_kernel(const int* __restrict__ lens, uint8_t* __restrict__ needed, int B, int F) {
  pdl_sync();
  const long long rows = static_cast<long long>(B) * F;
  const int tiles = static_cast<int>((rows + 127) / 128);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < tiles; t += gridDim.x * blockDim.x) {
    const long long r0 = static_cast<long long>(t) * 128, r1 = (r0 + 127 < rows - 1) ? r0 + 127 : rows - 1;
    const int b0 = static_cast<int>(r0 / F), b1 = static_cast<int>(r1 / F);
    int any = 0;
    for (int b = b0; b <= b1; ++b) {
      const long long first = b == b0 ? r0 - static_cast<long long>(b) * F : 0;     // first frame of utterance b inside the tile
      any |= first < lens[b] ? 1 : 0;
    }
    needed[t] = static_cast<uint8_t>(any);
  }
}

// f0[r] = gelu_tanh(z[r, :dp]) . w_f0 + b_f0, energy[r] = gelu_tanh(z[r, dp:]) . w_en + b_en; 0 past the frame count.
// One warp per frame row, DP = 128 * VPL columns per head, 128-bit loads.
template <int VPL>
__global__ void __launch_bounds__(256) prosody_head_kernel(const float* __restrict__ z, const float* __restrict__ w_f0,
                                                           const float* __restrict__ b_f0, const float* __restrict__ w_en,
                                                           const float* __restrict__ b_en, const int* __restrict__ lens,
                                                           float* __restrict__ f0, float* __restrict__ energy, int B, int F) {
  pdl_sync();
  constexpr int DP = 128 * VPL;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= static_cast<long long>(B) * F) return;
  const int b = static_cast<int>(row / F), f = static_cast<int>(row % F);
  if (f >= lens[b]) {                      // (rows of skipped GEMM tiles are never read)
    if (lane == 0) { f0[row] = 0.f; energy[row] = 0.f; }
    return;
  }
  const float4* zr = reinterpret_cast<const float4*>(z + row * 2 * DP);
  float acc[2] = {0.f, 0.f};
#pragma unroll
  for (int hd = 0; hd < 2; ++hd) {
    const float4* wr = reinterpret_cast<const float4*>(hd == 0 ? w_f0 : w_en);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float4 v = zr[hd * (DP >> 2) + i * 32 + lane], w = __ldg(wr + i * 32 + lane);
      acc[hd] += gelu_tanh(v.x) * w.x + gelu_tanh(v.y) * w.y + gelu_tanh(v.z) * w.z + gelu_tanh(v.w) * w.w;
    }
  }
  const float a0 = warp_sum(acc[0]), a1 = warp_sum(acc[1]);
  if (lane == 0) { f0[row] = a0 + __ldg(b_f0); energy[row] = a1 + __ldg(b_en); }
}

}  // namespace stz
