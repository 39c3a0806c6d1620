// Dense contractions of the style denoiser: C[M,N] = A[M,K] · W[N,K]^T (+ fused epilogue).
//
//   gemm2_kernel     — the product path (gemm2.cuh): persistent tcgen05 / TMEM / TMA kernel; this header holds the
//                      parameter block, the epilogue definitions and the small fp32 CUDA-core linears.
//   gemm_simt_kernel — same contract on CUDA cores: unit-test cross-check ("gemm_impl" = 1) and create-time M = 1 GEMMs.
//
// Row layout of denoiser activations ("R layout"): row r = (b*K + k)*2 + branch, branch 0 = cond,
// 1 = uncond, so the CFG pair of one style token sits in adjacent TMEM lanes of one tile and the
// guidance combine is a lane shuffle in the last GEMM's epilogue.
#pragma once
#include <cuda.h>
#include "ptx.cuh"
#include "elementwise.cuh"

namespace stz {

enum Epi : int {
  EPI_F32 = 0,        // out fp32 = acc + bias
  EPI_F32_POS = 1,    // out fp32 = acc + bias + pos[(r/2) % n_style]            (input projection)
  EPI_BF16 = 2,       // out bf16 = acc + bias                                    (QKV, Q, K/V)
  EPI_GELU_BF16 = 3,  // out bf16 = gelu_tanh(acc + bias)                         (FFN1)
  EPI_GATE_RES = 4,   // out fp32 += gate[seq(r)] * (acc + bias)                  (attn out, FFN2)
  EPI_SAMPLER = 5     // CFG combine + EDM/sampler affine update + next input     (output projection)
};

struct GemmParams {
  int M, N, K;          // M = valid rows of this problem
  int a_row0;           // first row inside A's tensor map
  const float* bias;    // [N] or nullptr
  void* out;            // [M, ldo] fp32 or bf16
  int ldo;
  // EPI_GATE_RES / EPI_F32_POS
  const float* mod;     // [n_seq, n_mod] AdaLN modulations of this eval
  int n_mod, gate_off;
  int rows_per_utt;     // rows of one utterance: 2 * n_style (CFG pair layout) or n_style (single-branch layout)
  const float* pos;     // [n_style, N]
  int n_style;
  int single;           // 0: R layout, row = (b*K + k)*2 + branch (CFG pair); 1: row = b*K + k (guidance-conditioned student)
  // EPI_SAMPLER: state x / xmid [B*K, N] fp32, noise slice [B*K, N], next input xin [M, N] bf16
  float* x;
  float* xmid;
  const float* noise;
  __nv_bfloat16* xin;
  const float* coef;    // device: {cx, cm, cF, cn, cin_next, cfg_scale, dest(0 = x, 1 = xmid), 0}
  float* tap;           // optional [B*K, N] copy of the guided F
  // optional [ceil(M / 128)] flags: 0 = every row of that 128-row tile is padding, the tile is skipped (its output rows
  // keep whatever they held; only legal when nothing downstream reads them: the duration predictor with text masks)
  const uint8_t* tile_needed;
};

// sequence (row of the modulation table) and style token of an activation row, for both row layouts
__device__ __forceinline__ int seq_of_row(int m, int rows_per_utt, int single) {
  return single ? m / rows_per_utt : (m / rows_per_utt) * 2 + (m & 1);
}
__device__ __forceinline__ int tok_of_row(int m, int n_style, int single) { return (single ? m : (m >> 1)) % n_style; }

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;

// One row (m), 32 consecutive columns starting at n0, accumulators in v[].  Called by all 32
// lanes of a warp whose lanes hold consecutive rows (needed by the EPI_SAMPLER pair shuffle).
template <int EPI>
__device__ __forceinline__ void epilogue_row32(const GemmParams& p, int m, int n0, float (&v)[32]) {
  const bool valid = m < p.M;
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if constexpr (EPI == EPI_F32 || EPI == EPI_F32_POS) {
    if (!valid) return;
    if constexpr (EPI == EPI_F32_POS) {
      const float* pr = p.pos + static_cast<size_t>(tok_of_row(m, p.n_style, p.single)) * p.N + n0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b = __ldg(reinterpret_cast<const float4*>(pr + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
    float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(m) * p.ldo + n0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  } else if constexpr (EPI == EPI_BF16 || EPI == EPI_GELU_BF16) {
    if (!valid) return;
    if constexpr (EPI == EPI_GELU_BF16) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(v[j]);
    }
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(m) * p.ldo + n0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      uint4 u;
      u.x = pack_bf16(v[j], v[j + 1]); u.y = pack_bf16(v[j + 2], v[j + 3]);
      u.z = pack_bf16(v[j + 4], v[j + 5]); u.w = pack_bf16(v[j + 6], v[j + 7]);
      *reinterpret_cast<uint4*>(o + j) = u;
    }
  } else if constexpr (EPI == EPI_GATE_RES) {
    if (!valid) return;
    const int seq = seq_of_row(m, p.rows_per_utt, p.single);
    const float* g = p.mod + static_cast<size_t>(seq) * p.n_mod + p.gate_off + n0;
    float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(m) * p.ldo + n0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 gg = __ldg(reinterpret_cast<const float4*>(g + j));
      float4 h = *reinterpret_cast<const float4*>(o + j);
      h.x += gg.x * v[j]; h.y += gg.y * v[j + 1]; h.z += gg.z * v[j + 2]; h.w += gg.w * v[j + 3];
      *reinterpret_cast<float4*>(o + j) = h;
    }
  } else {  // EPI_SAMPLER
    const float cx = __ldg(p.coef + 0), cm = __ldg(p.coef + 1), cF = __ldg(p.coef + 2), cn = __ldg(p.coef + 3);
    const float cin = __ldg(p.coef + 4), w = __ldg(p.coef + 5);
    const bool to_mid = __ldg(p.coef + 6) != 0.0f;
    const bool is_cond = p.single || (m & 1) == 0;
    if (!p.single) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float other = __shfl_xor_sync(0xffffffffu, v[j], 1);
        float Fc = is_cond ? v[j] : other, Fu = is_cond ? other : v[j];
        v[j] = Fu + w * (Fc - Fu);
      }
    }
    if (!valid) return;
    const size_t so = static_cast<size_t>(p.single ? m : (m >> 1)) * p.N + n0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 xv = *reinterpret_cast<const float4*>(p.x + so + j);
      float4 o = make_float4(cx * xv.x + cF * v[j], cx * xv.y + cF * v[j + 1], cx * xv.z + cF * v[j + 2],
                             cx * xv.w + cF * v[j + 3]);
      if (cm != 0.0f) {
        float4 xm = *reinterpret_cast<const float4*>(p.xmid + so + j);
        o.x += cm * xm.x; o.y += cm * xm.y; o.z += cm * xm.z; o.w += cm * xm.w;
      }
      if (cn != 0.0f) {
        float4 nz = __ldg(reinterpret_cast<const float4*>(p.noise + so + j));
        o.x += cn * nz.x; o.y += cn * nz.y; o.z += cn * nz.z; o.w += cn * nz.w;
      }
      if (is_cond) {
        *reinterpret_cast<float4*>((to_mid ? p.xmid : p.x) + so + j) = o;
        if (p.tap != nullptr) *reinterpret_cast<float4*>(p.tap + so + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      // next eval's input c_in(sigma') x' as the split-bf16 A operand [hi | lo | hi] (row stride 3N)
      const float4 y = make_float4(cin * o.x, cin * o.y, cin * o.z, cin * o.w);
      uint2 u;
      u.x = pack_bf16(y.x, y.y); u.y = pack_bf16(y.z, y.w);
      __nv_bfloat16* xo = p.xin + static_cast<size_t>(m) * 3 * p.N + n0 + j;
      *reinterpret_cast<uint2*>(xo) = u;
      *reinterpret_cast<uint2*>(xo + p.N) = split_lo4(y, u);
      *reinterpret_cast<uint2*>(xo + 2 * p.N) = u;
    }
  }
}


// CUDA-core cross-check of the same contract (block = 128 rows x 32 columns, thread = row).
template <int EPI>
__global__ void __launch_bounds__(128) gemm_simt_kernel(const __nv_bfloat16* __restrict__ A, int lda,
                                                        const __nv_bfloat16* __restrict__ W, const GemmParams p) {
  __shared__ float ws[32][65];
  const int n0 = blockIdx.x * 32;
  const int m = blockIdx.y * 128 + threadIdx.x;
  const bool valid = m < p.M;
  const __nv_bfloat16* a = A + static_cast<size_t>(p.a_row0 + (valid ? m : 0)) * lda;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = 0.f;
  for (int k0 = 0; k0 < p.K; k0 += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 64; i += 128) {
      int j = i >> 6, k = i & 63;
      ws[j][k] = __bfloat162float(W[static_cast<size_t>(n0 + j) * p.K + k0 + k]);
    }
    __syncthreads();
    for (int k = 0; k < 64; ++k) {
      float av = __bfloat162float(a[k0 + k]);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaf(av, ws[j][k], v[j]);
    }
  }
  epilogue_row32<EPI>(p, m, n0, v);
}

// ------------------------------------------------------------------------------------------
// fp32 CUDA-core GEMM for the duration predictor (true fp32: SURVEY.md §7 hard part 3) and the
// small conditioning projections:  Y[M,N] = act([X1 | X2][M, K1+K2] · W[N, K1+K2]^T + b).
// 64x64 tile, BK = 16, 256 threads, 4x4 outputs per thread.
// ------------------------------------------------------------------------------------------
enum Act : int { ACT_NONE = 0, ACT_SILU = 1 };

template <int ACT>
__global__ void __launch_bounds__(256) linear_f32_kernel(const float* __restrict__ X1, int ld1, int K1,
                                                         const float* __restrict__ X2, int ld2, int K2,
                                                         const float* __restrict__ W, const float* __restrict__ b,
                                                         float* __restrict__ Y, int ldy, int M, int N) {
  pdl_sync();
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float xs[BK][BM + 4];
  __shared__ float wsm[BK][BN + 4];
  const int K = K1 + K2;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += BK) {
    // 64 rows x 16 k = 1024 elements per operand, 4 per thread; consecutive threads -> consecutive k
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + i * 256;
      const int r = idx >> 4, kk = idx & 15, k = k0 + kk;
      float xv = 0.f, wv = 0.f;
      if (k < K) {
        const int m = m0 + r, n = n0 + r;
        if (m < M) xv = (k < K1) ? X1[static_cast<size_t>(m) * ld1 + k] : X2[static_cast<size_t>(m) * ld2 + (k - K1)];
        if (n < N) wv = W[static_cast<size_t>(n) * K + k];
      }
      xs[kk][r] = xv;
      wsm[kk][r] = wv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&xs[kk][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&wsm[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float y = acc[i][j] + (b != nullptr ? b[n] : 0.f);
      if (ACT == ACT_SILU) y = silu(y);
      Y[static_cast<size_t>(m) * ldy + n] = y;
    }
  }
}

// Small-M fp32 linear (pooled-conditioning projections, time-embedding MLP: M <= a few hundred rows).
// One warp per (row m, 8 consecutive outputs): lanes split K with coalesced 128-bit loads, 8 independent
// accumulators, one batched warp reduction.  K % 128 == 0, N % 8 == 0.
template <int ACT>
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                                          const float* __restrict__ b, float* __restrict__ Y, int ldy,
                                                          int M, int N, int K) {
  pdl_sync();
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int ngrp = N >> 3;
  if (warp >= M * ngrp) return;
  const int m = warp / ngrp, n0 = (warp % ngrp) << 3;
  const float4* xr = reinterpret_cast<const float4*>(X + static_cast<size_t>(m) * ldx);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int k4 = lane; k4 < (K >> 2); k4 += 32) {
    const float4 x = xr[k4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(W + static_cast<size_t>(n0 + j) * K) + k4);
      acc[j] = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, acc[j]))));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  if (lane < 8) {
    float y = acc[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) y = lane == j ? acc[j] : y;
    y += b != nullptr ? b[n0 + lane] : 0.f;
    if (ACT == ACT_SILU) y = silu(y);
    Y[static_cast<size_t>(m) * ldy + n0 + lane] = y;
  }
}

}  // namespace stz
