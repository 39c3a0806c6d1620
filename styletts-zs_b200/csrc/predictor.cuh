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
                                                              const float* __restrict__ v, int ldkv, float* __restrict__ out,
                                                              int n_tok, int T, int K, int ds, float scale) {
  pdl_sync();
  const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= n_tok) return;
  const int b = tok / T;
  const int nh = ds >> 5;
  for (int h = 0; h < nh; ++h) {
    const float qv = q[static_cast<size_t>(tok) * ds + h * 32 + lane] * scale;
    float m = -INFINITY, l = 0.f, acc = 0.f;
    for (int j = 0; j < K; ++j) {
      const size_t r = (static_cast<size_t>(b) * K + j) * ldkv + h * 32 + lane;
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

// needed[t] = 1 iff any of the rows 128 t .. 128 t + 127 of a [rows] mask is valid: the predictor's GEMMs skip 128-row
// tiles that hold nothing but padding (variable-length batches: about half of cfg4's B x T rows).
// tiles_per_utt > 0 (utterances aligned to tiles): the FIRST tile of every utterance is always needed, so consumers that
// unconditionally visit an utterance's first key block (attention_tcs_kernel) read initialised, finite rows even for a
// zero-length utterance.
__global__ void __launch_bounds__(256) tile_needed_kernel(const uint8_t* __restrict__ mask, uint8_t* __restrict__ needed, int rows,
                                                          int tiles_per_utt) {
  pdl_sync();
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int r0 = warp * 128;
  if (r0 >= rows) return;
  int any = (tiles_per_utt > 0 && warp % tiles_per_utt == 0) ? 1 : 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + i * 32 + lane;
    any |= (r < rows && mask[r] != 0) ? 1 : 0;
  }
  any = __any_sync(0xffffffffu, any);
  if (lane == 0) needed[warp] = static_cast<uint8_t>(any);
}

// a-8, product kernel: one CTA = SP_TOK consecutive tokens of one utterance; the utterance's projected K / V
// (K x ds fp32 each) are staged in shared memory once, then a warp handles one token at a time:
//   scores  : lane = key (keys j and j + 32), q broadcast by shuffle, K rows padded to ds + 1 (conflict-free);
//   softmax : two warp reductions per (token, head) instead of one per key;
//   output  : lane = channel, p_j broadcast by shuffle, V rows read conflict-free.
constexpr int SP_TOK = 16;
__global__ void __launch_bounds__(256) style_pool_attn2_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                               const float* __restrict__ v, int ldkv, float* __restrict__ out,
                                                               int T, int K, int ds, float scale) {
  extern __shared__ float sp_smem[];
  float* Ks = sp_smem;                  // [K][ds + 1]
  float* Vs = sp_smem + ((K * (ds + 1) + 3) & ~3);   // [K][ds], 16-byte aligned
  pdl_sync();
  const int b = blockIdx.y, t0 = blockIdx.x * SP_TOK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < K * (ds >> 2); i += blockDim.x) {
    const int j = i / (ds >> 2), c = (i % (ds >> 2)) * 4;
    const size_t r = (static_cast<size_t>(b) * K + j) * ldkv + c;
    const float4 kk = __ldg(reinterpret_cast<const float4*>(k + r)), vv = __ldg(reinterpret_cast<const float4*>(v + r));
    float* kd = Ks + j * (ds + 1) + c;
    kd[0] = kk.x; kd[1] = kk.y; kd[2] = kk.z; kd[3] = kk.w;
    *reinterpret_cast<float4*>(Vs + j * ds + c) = vv;
  }
  __syncthreads();
  const int nh = ds >> 5;
  const int j0 = lane < K ? lane : K - 1, j1 = lane + 32 < K ? lane + 32 : K - 1;
  for (int tt = warp; tt < SP_TOK && t0 + tt < T; tt += blockDim.x >> 5) {
    const size_t tok = static_cast<size_t>(b) * T + t0 + tt;
    for (int h = 0; h < nh; ++h) {
      const float qv = q[tok * ds + h * 32 + lane] * scale;
      const float* k0 = Ks + j0 * (ds + 1) + h * 32;
      const float* k1 = Ks + j1 * (ds + 1) + h * 32;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float qc = __shfl_sync(0xffffffffu, qv, c);
        s0 = fmaf(qc, k0[c], s0);
        s1 = fmaf(qc, k1[c], s1);
      }
      if (lane >= K) s0 = -INFINITY;
      if (lane + 32 >= K) s1 = -INFINITY;
      float m = fmaxf(s0, s1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      const float p0 = expf(s0 - m), p1 = expf(s1 - m);   // exp(-inf) = 0 for the slots past K
      const float l = warp_sum(p0 + p1);
      float acc = 0.f;
      const float* vc = Vs + h * 32 + lane;
      for (int j = 0; j < K; ++j) {
        const float pj = __shfl_sync(0xffffffffu, j < 32 ? p0 : p1, j & 31);
        acc = fmaf(pj, vc[j * ds], acc);
      }
      out[tok * ds + h * 32 + lane] = acc / l;
    }
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
                                                         const uint8_t* __restrict__ mask, int rows,
                                                         __nv_bfloat16* __restrict__ a3, int ldd, int segK) {
  pdl_sync();
  constexpr int D = 128 * VPL;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float4* xr = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * D);
  if (mask != nullptr && mask[row] == 0) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      xr[i * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a3 != nullptr) {
        __nv_bfloat16* o = a3 + static_cast<size_t>(row) * ldd + (i * 32 + lane) * 4;
        *reinterpret_cast<uint2*>(o) = make_uint2(0, 0);
        *reinterpret_cast<uint2*>(o + segK) = make_uint2(0, 0);
        *reinterpret_cast<uint2*>(o + 2 * segK) = make_uint2(0, 0);
      }
    }
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
    const float4 y = make_float4(v[i].x * rstd * (1.f + ga.x) + be.x, v[i].y * rstd * (1.f + ga.y) + be.y,
                                 v[i].z * rstd * (1.f + ga.z) + be.z, v[i].w * rstd * (1.f + ga.w) + be.w);
    xr[i * 32 + lane] = y;
    if (a3 != nullptr) {  // next layer's split-bf16 GEMM operand (x part of [x | s_tok])
      uint2 hi;
      hi.x = pack_bf16(y.x, y.y); hi.y = pack_bf16(y.z, y.w);
      __nv_bfloat16* o = a3 + static_cast<size_t>(row) * ldd + (i * 32 + lane) * 4;
      *reinterpret_cast<uint2*>(o) = hi;
      *reinterpret_cast<uint2*>(o + segK) = split_lo4(y, hi);
      *reinterpret_cast<uint2*>(o + 2 * segK) = hi;
    }
  }
}

// a-10: dur = clamp(rint(sum_j sigmoid(x·Wd_j + b_j)), min 1) * mask  -> int32 on device.
// rintf == round-half-to-even == torch.round.
template <int VPL>
__global__ void __launch_bounds__(256) dur_head_kernel(const float* __restrict__ x, const float* __restrict__ Wd,
                                                       const float* __restrict__ bd, const uint8_t* __restrict__ mask,
                                                       int32_t* __restrict__ dur, float* __restrict__ presum, int rows,
                                                       int max_dur) {
  pdl_sync();
  constexpr int D = 128 * VPL;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
  float4 v[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) v[i] = xr[i * 32 + lane];
  // 10 logits at a time: independent partial dot products, then one batched warp reduction (no serial latency chain)
  float total = 0.f;
  for (int j0 = 0; j0 < max_dur; j0 += 10) {
    float d[10];
#pragma unroll
    for (int jj = 0; jj < 10; ++jj) {
      d[jj] = 0.f;
      if (j0 + jj < max_dur) {
        const float4* wr = reinterpret_cast<const float4*>(Wd + static_cast<size_t>(j0 + jj) * D);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const float4 w = __ldg(wr + i * 32 + lane);
          d[jj] += v[i].x * w.x + v[i].y * w.y + v[i].z * w.z + v[i].w * w.w;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int jj = 0; jj < 10; ++jj) d[jj] += __shfl_xor_sync(0xffffffffu, d[jj], o);
#pragma unroll
    for (int jj = 0; jj < 10; ++jj)
      if (j0 + jj < max_dur) total += sigmoidf_(d[jj] + __ldg(bd + j0 + jj));
  }
  if (lane == 0) {
    const bool ok = mask == nullptr || mask[row] != 0;
    if (presum != nullptr) presum[row] = total;
    dur[row] = ok ? static_cast<int32_t>(fmaxf(rintf(total), 1.0f)) : 0;
  }
}


// a-10, product kernel: one CTA = 32 tokens staged in shared memory; warp w evaluates logits w, w + 8, ... against
// all 32 tokens (the logit's weight row lives in registers: W is read once per CTA, not once per token), the 32
// per-token dot products of a logit are reduced together with a 31-shuffle transpose reduction (lane r ends up
// with token r), and the per-warp sigmoid sums are combined in a fixed order (deterministic).
template <int VPL>
__global__ void __launch_bounds__(256) dur_head2_kernel(const float* __restrict__ x, const float* __restrict__ Wd,
                                                        const float* __restrict__ bd, const uint8_t* __restrict__ mask,
                                                        int32_t* __restrict__ dur, float* __restrict__ presum, int rows,
                                                        int max_dur) {
  constexpr int D = 128 * VPL;
  extern __shared__ __align__(16) float dh_smem[];
  float4* xs = reinterpret_cast<float4*>(dh_smem);                 // [32][D / 4]
  float* part = dh_smem + 32 * D;                                   // [8][32]
  pdl_sync();
  const int row0 = blockIdx.x * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32 * (D >> 2); i += blockDim.x) {
    const int r = i / (D >> 2);
    xs[i] = row0 + r < rows ? __ldg(reinterpret_cast<const float4*>(x + static_cast<size_t>(row0) * D) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  float total = 0.f;   // lane r: sum over this warp's logits of sigmoid(logit) for token r
  for (int j = warp; j < max_dur; j += 8) {
    float4 w[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) w[i] = __ldg(reinterpret_cast<const float4*>(Wd + static_cast<size_t>(j) * D) + i * 32 + lane);
    float d[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const float4 xv = xs[r * (D >> 2) + i * 32 + lane];
        a = fmaf(xv.x, w[i].x, fmaf(xv.y, w[i].y, fmaf(xv.z, w[i].z, fmaf(xv.w, w[i].w, a))));
      }
      d[r] = a;
    }
    // transpose reduction: after the step with offset o, lane keeps the rows whose bit o equals its own bit o
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int r = 0; r < o; ++r) {
        const float keep = up ? d[r + o] : d[r], give = up ? d[r] : d[r + o];
        d[r] = keep + __shfl_xor_sync(0xffffffffu, give, o);
      }
    }
    total += sigmoidf_(d[0] + __ldg(bd + j));   // d[0] now holds the full dot product of token `lane`
  }
  part[warp * 32 + lane] = total;
  __syncthreads();
  if (warp == 0) {
    const int row = row0 + lane;
    if (row < rows) {
      float t = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) t += part[w8 * 32 + lane];
      const bool ok = mask == nullptr || mask[row] != 0;
      if (presum != nullptr) presum[row] = t;
      dur[row] = ok ? static_cast<int32_t>(fmaxf(rintf(t), 1.0f)) : 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// a-9, product path: persistent thread-block-cluster BiLSTM recurrence (h = 256).
//
// One cluster of 8 CTAs advances NB sequences of one direction through all their time steps.
// CTA r owns hidden units [32r, 32r+32) = 128 gate columns; its 128 x 256 fp32 slice of W_hh lives
// in REGISTERS for the whole kernel (512 threads x 64 values: thread = (gate column, k-quarter)), so a
// step costs no weight traffic at all.  h_t of the NB sequences is replicated in every CTA's shared
// memory (double-buffered); after the pointwise update each CTA pushes its 32 new units to the 7 peers
// with st.shared::cluster (DSMEM) and one barrier.cluster per step orders the exchange.
// Sequences are visited in length-sorted order (perm) so the NB sequences of a cluster have similar
// lengths; packed-sequence semantics as lstm_rec_kernel (reverse starts at len-1, padded outputs 0).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t local_smem_addr, uint32_t rank, float v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

// lens[b] (prefix-mask popcount) and perm = sequence indices sorted by length, longest first (stable).
__global__ void __launch_bounds__(1024) lens_perm_kernel(const uint8_t* __restrict__ mask, int* __restrict__ lens,
                                                         int* __restrict__ perm, int B, int T) {
  pdl_sync();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int n = T;
    if (mask != nullptr) {
      n = 0;
      for (int t = 0; t < T; ++t) n += mask[static_cast<size_t>(b) * T + t] ? 1 : 0;
    }
    lens[b] = n;
  }
  __syncthreads();
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

constexpr int LC_H = 256, LC_CS = 8, LC_UPC = 32, LC_COLS = 128, LC_SL = 8, LC_KS = 32, LC_THREADS = 512;

// DSMEM store that signals the destination CTA's mbarrier with the bytes it delivered (no fence, no cluster barrier)
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];"
               ::"r"(remote_addr), "f"(v), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(rank));
  return remote;
}
// CTA-scope wait: st.async data is written through the async proxy and published by the barrier's complete_tx,
// exactly like a TMA load, so no cluster-scope acquire (which would invalidate L1 every step) is needed.
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

}  // namespace stz
