// Counter-based Gaussian noise on the device: Philox4x32-10 + Box-Muller whose log / sin / cos are
// fixed polynomial evaluations of individually rounded fp32 operations (no FMA contraction, no libm), so the numpy
// restatement in oracle/philox.py reproduces every output bit.  key = seed; counter = (group of 4 elements inside the
// utterance's [n_per_utt] slice, utterance lo, noise slice, utterance hi) — an utterance's noise does not depend on the
// batch or the GPU it is sampled on.  HBM-bound writer: 16 B per thread per counter, coalesced float4 stores.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"

namespace philox {

__device__ __forceinline__ void round4x32(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  c[0] = hi1 ^ c[1] ^ k0;
  c[1] = lo1;
  c[2] = hi0 ^ c[3] ^ k1;
  c[3] = lo0;
}

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    round4x32(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// every step is a separately rounded fp32 operation: the *_rn intrinsics are never contracted into FMAs
__device__ __forceinline__ float mad_rn(float a, float b, float c) { return __fadd_rn(__fmul_rn(a, b), c); }

__device__ __forceinline__ float log_u(float u) {   // ln(u), u in (0, 1)
  const uint32_t bits = __float_as_uint(u);
  int e = static_cast<int>(bits >> 23) - 127;
  float m = __uint_as_float((bits & 0x007FFFFFu) | 0x3F800000u);
  if (m > 1.41421354f) { m = __fmul_rn(m, 0.5f); e += 1; }
  const float s = __fdiv_rn(__fadd_rn(m, -1.0f), __fadd_rn(m, 1.0f));
  const float s2 = __fmul_rn(s, s);
  float p = 0.111111112f;
  p = mad_rn(p, s2, 0.142857149f);
  p = mad_rn(p, s2, 0.200000003f);
  p = mad_rn(p, s2, 0.333333343f);
  p = mad_rn(p, s2, 1.0f);
  const float lnm = __fmul_rn(__fmul_rn(2.0f, s), p);
  return __fadd_rn(__fmul_rn(static_cast<float>(e), 0.693147182f), lnm);
}

__device__ __forceinline__ void sincos_turn(uint32_t k, float& s, float& c) {   // angle 2 pi (k + 0.5) / 2^23
  const uint32_t q = k >> 21;
  uint32_t j = k & 0x1FFFFFu;
  const bool swap = j >= (1u << 20);
  if (swap) j = ((1u << 21) - 1u) - j;
  const float phi = __fmul_rn(__fadd_rn(static_cast<float>(j), 0.5f), 7.49014077e-07f);
  const float x2 = __fmul_rn(phi, phi);
  float ps = 2.75573188e-06f;
  ps = mad_rn(ps, x2, -1.98412701e-04f);
  ps = mad_rn(ps, x2, 8.33333377e-03f);
  ps = mad_rn(ps, x2, -1.66666672e-01f);
  const float sn = __fmul_rn(phi, mad_rn(ps, x2, 1.0f));
  float pc = -2.75573188e-07f;
  pc = mad_rn(pc, x2, 2.48015876e-05f);
  pc = mad_rn(pc, x2, -1.38888892e-03f);
  pc = mad_rn(pc, x2, 4.16666679e-02f);
  pc = mad_rn(pc, x2, -0.5f);
  const float cs = mad_rn(pc, x2, 1.0f);
  const float sq = swap ? cs : sn, cq = swap ? sn : cs;
  s = q == 0 ? sq : q == 1 ? cq : q == 2 ? -sq : -cq;
  c = q == 0 ? cq : q == 1 ? -sq : q == 2 ? -cq : sq;
}

__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& z0, float& z1) {
  const float u = __fmul_rn(__fadd_rn(static_cast<float>(xa >> 9), 0.5f), 1.1920929e-07f);   // 2^-23
  const float r = __fsqrt_rn(__fmul_rn(-2.0f, log_u(u)));
  float s, c;
  sincos_turn(xb >> 9, s, c);
  z0 = __fmul_rn(r, c);
  z1 = __fmul_rn(r, s);
}

}  // namespace philox

// out [slices, B, n_per_utt] fp32, n_per_utt % 4 == 0; one thread per group of four elements; utterance b's global index is
// utt_ids[b] when given (sharded batches: any subset, any order), else first_utt + b
__global__ void __launch_bounds__(256) philox_normal_kernel(float* __restrict__ out, uint32_t seed_lo, uint32_t seed_hi,
                                                            unsigned long long first_utt,
                                                            const unsigned long long* __restrict__ utt_ids, int slices,
                                                            int B, int groups) {
  stz::pdl_sync();   // `out` may still be read by the previous call's kernels
  const size_t total = static_cast<size_t>(slices) * B * groups;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint32_t g = static_cast<uint32_t>(i % groups);
    const size_t sb = i / groups;
    const unsigned long long utt = utt_ids ? utt_ids[sb % B] : first_utt + sb % B;
    uint32_t c[4] = {g, static_cast<uint32_t>(utt), static_cast<uint32_t>(sb / B), static_cast<uint32_t>(utt >> 32)};
    philox::philox4x32_10(c, seed_lo, seed_hi);
    float4 z;
    philox::box_muller(c[0], c[1], z.x, z.y);
    philox::box_muller(c[2], c[3], z.z, z.w);
    reinterpret_cast<float4*>(out)[i] = z;
  }
}
