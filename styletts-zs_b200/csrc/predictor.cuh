// Duration predictor kernels (SURVEY.md §8 a-8 .. a-10).  Everything here is true fp32 on CUDA
// cores: integer durations must match the oracle on >= 99.9 % of tokens, and a bf16/TF32 error on
// the 0..50 sigmoid sum would flip ~1 % of the roundings (SURVEY.md §7 hard part 3).
#pragma once
#include "elementwise.cuh"

namespace stz {

// lens[b] = number of valid tokens (prefix mask) or T when mask == nullptr
__global__ void lens_kernel(const uint8_t* __restrict__ mask, int* __restrict__ lens, int B, int T) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int n = T;
  if (mask != nullptr) {
    n = 0;
    for (int t = 0; t < T; ++t) n += mask[static_cast<size_t>(b) * T + t] ? 1 : 0;
  }
  lens[b] = n;
}

// a-8: per-token style summary.  q [B*T, ds], k/v [B*K, ds] (already projected), out [B*T, ds];
// heads of width 32 (lane = channel).  One warp per token, online softmax over the K style codes.
__global__ void __launch_bounds__(256) style_pool_attn_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                              const float* __restrict__ v, float* __restrict__ out,
                                                              int n_tok, int T, int K, int ds, float scale) {
  const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= n_tok) return;
  const int b = tok / T;
  const int nh = ds >> 5;
  for (int h = 0; h < nh; ++h) {
    const float qv = q[static_cast<size_t>(tok) * ds + h * 32 + lane] * scale;
    float m = -INFINITY, l = 0.f, acc = 0.f;
    for (int j = 0; j < K; ++j) {
      const size_t r = (static_cast<size_t>(b) * K + j) * ds + h * 32 + lane;
      const float s = warp_sum(qv * __ldg(k + r));
      const float mn = fmaxf(m, s);
      const float a = expf(m - mn), pj = expf(s - mn);
      l = l * a + pj;
      acc = acc * a + pj * __ldg(v + r);
      m = mn;
    }
    out[static_cast<size_t>(tok) * ds + h * 32 + lane] = acc / l;
  }
}

// a-9: one direction of one BiLSTM layer for a chunk of NB sequences (packed-sequence semantics:
// the reverse direction starts at each sequence's own last valid token; padded outputs are 0).
//   G    [B*T, 8h]   W_ih·[x, s_tok] + b_ih + b_hh for both directions, column = dir*4h + gate*h + unit
//   WhhT [2, h, 4h]  recurrent weights, k-major (transposed at create time) so lanes read coalesced
//   out  [B, T, 2h]  forward half in columns [0,h), reverse half in [h,2h)
// grid = (ceil(B/NB), 2), block = 4h threads (one gate column each).
template <int NB>
__global__ void __launch_bounds__(1024) lstm_rec_kernel(const float* __restrict__ G, const float* __restrict__ WhhT,
                                                        const int* __restrict__ lens, float* __restrict__ out, int B,
                                                        int T, int h) {
  extern __shared__ float sm[];
  float* h_s = sm;                 // [NB][h]
  float* g_s = sm + NB * h;        // [NB][4h]
  __shared__ int len_s[NB];
  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * NB;
  const int col = threadIdx.x;     // gate*h + unit
  const int H4 = 4 * h;
  if (threadIdx.x < NB) len_s[threadIdx.x] = (b0 + threadIdx.x < B) ? lens[b0 + threadIdx.x] : 0;
  for (int i = threadIdx.x; i < NB * h; i += blockDim.x) h_s[i] = 0.f;
  __syncthreads();
  int maxlen = 0;
#pragma unroll
  for (int n = 0; n < NB; ++n) maxlen = max(maxlen, len_s[n]);
  // zero the padded tail of this direction's half of the output
  for (int n = 0; n < NB; ++n) {
    if (b0 + n >= B) break;
    const int tail = (T - len_s[n]) * h;
    float* o = out + (static_cast<size_t>(b0 + n) * T + len_s[n]) * 2 * h + dir * h;
    for (int i = threadIdx.x; i < tail; i += blockDim.x) o[static_cast<size_t>(i / h) * 2 * h + (i % h)] = 0.f;
  }
  // pointwise phase ownership: thread -> (sequence slot n4 + 4*i, unit)
  const int unit = threadIdx.x % h, n4 = threadIdx.x / h;
  float c_state[NB / 4];
#pragma unroll
  for (int i = 0; i < NB / 4; ++i) c_state[i] = 0.f;
  const float* W = WhhT + static_cast<size_t>(dir) * h * H4 + col;

  for (int s = 0; s < maxlen; ++s) {
    float acc[NB];
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      acc[n] = 0.f;
      if (s < len_s[n]) {
        const int t = dir == 0 ? s : len_s[n] - 1 - s;
        acc[n] = __ldg(G + (static_cast<size_t>(b0 + n) * T + t) * 2 * H4 + dir * H4 + col);
      }
    }
    for (int k = 0; k < h; k += 4) {
      const float w0 = __ldg(W + static_cast<size_t>(k) * H4), w1 = __ldg(W + static_cast<size_t>(k + 1) * H4);
      const float w2 = __ldg(W + static_cast<size_t>(k + 2) * H4), w3 = __ldg(W + static_cast<size_t>(k + 3) * H4);
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        const float4 hv = *reinterpret_cast<const float4*>(h_s + n * h + k);
        acc[n] = fmaf(w0, hv.x, acc[n]);
        acc[n] = fmaf(w1, hv.y, acc[n]);
        acc[n] = fmaf(w2, hv.z, acc[n]);
        acc[n] = fmaf(w3, hv.w, acc[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < NB; ++n) g_s[n * H4 + col] = acc[n];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NB / 4; ++i) {
      const int n = n4 + 4 * i;
      if (s < len_s[n]) {
        const float* gg = g_s + n * H4 + unit;
        const float ig = sigmoidf_(gg[0]), fg = sigmoidf_(gg[h]), gt = tanhf(gg[2 * h]), og = sigmoidf_(gg[3 * h]);
        const float c = fg * c_state[i] + ig * gt;
        c_state[i] = c;
        const float hn = og * tanhf(c);
        h_s[n * h + unit] = hn;
        const int t = dir == 0 ? s : len_s[n] - 1 - s;
        out[(static_cast<size_t>(b0 + n) * T + t) * 2 * h + dir * h + unit] = hn;
      }
    }
    __syncthreads();
  }
}

// a-9: x[r] = (LN(x[r]) * (1 + gamma[r]) + beta[r]) * mask[r], gb[r] = [gamma | beta].  In place.
template <int VPL>
__global__ void __launch_bounds__(256) adaln_pred_kernel(float* __restrict__ x, const float* __restrict__ gb,
                                                         const uint8_t* __restrict__ mask, int rows) {
  constexpr int D = 128 * VPL;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float4* xr = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * D);
  if (mask != nullptr && mask[row] == 0) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) xr[i * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = xr[i * 32 + lane];
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
  const float4* gr = reinterpret_cast<const float4*>(gb + static_cast<size_t>(row) * 2 * D);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float4 ga = __ldg(gr + i * 32 + lane), be = __ldg(gr + (D >> 2) + i * 32 + lane);
    xr[i * 32 + lane] = make_float4(v[i].x * rstd * (1.f + ga.x) + be.x, v[i].y * rstd * (1.f + ga.y) + be.y,
                                    v[i].z * rstd * (1.f + ga.z) + be.z, v[i].w * rstd * (1.f + ga.w) + be.w);
  }
}

// a-10: dur = clamp(rint(sum_j sigmoid(x·Wd_j + b_j)), min 1) * mask  -> int32 on device.
// rintf == round-half-to-even == torch.round.
template <int VPL>
__global__ void __launch_bounds__(256) dur_head_kernel(const float* __restrict__ x, const float* __restrict__ Wd,
                                                       const float* __restrict__ bd, const uint8_t* __restrict__ mask,
                                                       int32_t* __restrict__ dur, float* __restrict__ presum, int rows,
                                                       int max_dur) {
  constexpr int D = 128 * VPL;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
  float4 v[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) v[i] = xr[i * 32 + lane];
  float total = 0.f;
  for (int j = 0; j < max_dur; ++j) {
    const float4* wr = reinterpret_cast<const float4*>(Wd + static_cast<size_t>(j) * D);
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float4 w = __ldg(wr + i * 32 + lane);
      d += v[i].x * w.x + v[i].y * w.y + v[i].z * w.z + v[i].w * w.w;
    }
    d = warp_sum(d) + __ldg(bd + j);
    total += sigmoidf_(d);
  }
  if (lane == 0) {
    const bool ok = mask == nullptr || mask[row] != 0;
    if (presum != nullptr) presum[row] = total;
    dur[row] = ok ? static_cast<int32_t>(fmaxf(rintf(total), 1.0f)) : 0;
  }
}

}  // namespace stz
