// gemmln_kernel — a residual-writing GEMM fused with the AdaLN that follows it (style denoiser, N = d_model = 512).
//
// One CTA owns a full 128-row x 512-column block of the residual stream, so the row statistics of the
// LayerNorm that the next sub-layer needs are available in the epilogue:
//     acc   = A[128, K] · W[512, K]^T                                  tcgen05.mma, 4 x (128 x 128) fp32 in TMEM (512 columns)
//     h'    = h + gate[seq] * (acc + b)        (MODE_RES)      or      acc + b + pos[(r/2) % n_style]   (MODE_POS)
//     u     = bf16( LN(h') * (1 + scale[seq]) + shift[seq] )           the next GEMM's A operand ([hi|lo|hi] if split3)
// This replaces {GEMM with gated residual epilogue, ln_mod_kernel}: one launch instead of two, h' written once,
// u produced without re-reading h'.
//
// Warp roles: warp 0 TMA producer (A ring: 3 x 16 KB; W ring: 4 x 16 KB quarter tiles of 128 rows x 64 k),
// warp 1 MMA issuer (k outer, column quarter inner), warps 2..9 epilogue: thread = row (TMEM lane), two warps per
// lane quadrant split the 512 columns in halves.  Pass 1: residual tiles TMA-loaded into staging, h' added in place,
// TMA-stored and kept in TMEM (tcgen05.st),
// partial sum / sum of squares; halves exchange statistics through shared memory; pass 2: u -> swizzled staging ->
// TMA store.
#pragma once
#include "gemm2.cuh"

namespace stz {

enum GlnMode : int { GLN_RES = 0, GLN_POS = 1 };

struct GemmLnParams {
  int M, K;               // valid rows, contraction length (multiple of 64)
  const float* bias;      // [512]
  float* h;               // [M, 512] residual stream (read in GLN_RES, always written)
  const float* mod;       // [n_seq, n_mod] AdaLN modulations of this evaluation
  int n_mod, gate_off, shift_off, scale_off;
  int rows_per_utt;       // 2 * n_style
  const float* pos;       // [n_style, 512]  (GLN_POS)
  int n_style;
  int split3;             // u is the split-bf16 operand [hi | lo | hi] with row stride 3 * 512
};

constexpr int GLN_N = 512, GLN_SA = 3, GLN_SB = 4, GLN_THREADS = 320;
constexpr int GLN_A_BYTES = GEMM_BM * GEMM_BK * 2, GLN_B_BYTES = 128 * GEMM_BK * 2;
constexpr int GLN_SMEM_BYTES = GLN_SA * GLN_A_BYTES + GLN_SB * GLN_B_BYTES + 8 * 3 * 4096 + 1024;   // + 3 staging tiles per epilogue warp

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(GLN_THREADS, 1) gemmln_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmU,
                                                               const __grid_constant__ CUtensorMap tmH,
                                                               const GemmLnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[GLN_SA], a_empty[GLN_SA], b_full[GLN_SB], b_empty[GLN_SB], acc_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float2 stats_s[2][GEMM_BM];
  __shared__ __align__(8) uint64_t hbar[8][3];   // per epilogue warp: residual chunk landed in staging tile i

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base + GLN_SA * GLN_A_BYTES;
  const uint32_t epi_base = b_base + GLN_SB * GLN_B_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = p.K / GEMM_BK;
  const int tile_m = blockIdx.x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmU);
    prefetch_tmap(&tmH);
    for (int s = 0; s < GLN_SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < GLN_SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(&acc_full, 1);
    for (int i = 0; i < 24; ++i) mbar_init(&hbar[0][0] + i, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_sync();

  if (warp == 0) {
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&a_empty[sa], pa ^ 1u);
        mbar_expect_tx(&a_full[sa], GLN_A_BYTES);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(smem_base + sa * GLN_A_BYTES), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(smem_u32(&a_full[sa])),
            "r"(kb * GEMM_BK), "r"(tile_m * GEMM_BM)
            : "memory");
        if (++sa == GLN_SA) { sa = 0; pa ^= 1u; }
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          mbar_wait(&b_empty[sb], pb ^ 1u);
          mbar_expect_tx(&b_full[sb], GLN_B_BYTES);
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
              ::"r"(b_base + sb * GLN_B_BYTES), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(&b_full[sb])),
              "r"(kb * GEMM_BK), "r"(q * 128)
              : "memory");
          if (++sb == GLN_SB) { sb = 0; pb ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, 128);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&a_full[sa], pa);
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(smem_base + sa * GLN_A_BYTES);
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          mbar_wait(&b_full[sb], pb);
          tc_fence_after();
          const uint64_t db = umma_desc_sw128(b_base + sb * GLN_B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_bf16(tmem_base + q * 128, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&b_empty[sb]);
          if (++sb == GLN_SB) { sb = 0; pb ^= 1u; }
        }
        umma_commit(&a_empty[sa]);
        if (++sa == GLN_SA) { sa = 0; pa ^= 1u; }
      }
      umma_commit(&acc_full);
    }
  } else {
    const int q4 = warp & 3, half = (warp - 2) >> 2;
    const int r_in = q4 * 32 + lane;                // row inside the tile == TMEM lane
    const int m = tile_m * GEMM_BM + r_in;
    const int mm = m < p.M ? m : p.M - 1;           // clamp for loads; rows >= M are never stored
    const float* mrow = p.mod + static_cast<size_t>((mm / p.rows_per_utt) * 2 + (mm & 1)) * p.n_mod;
    const float* src = p.pos + static_cast<size_t>((mm >> 1) % p.n_style) * GLN_N;   // GLN_POS only
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + half * 256;
    const int col0 = half * 256;

    // ---- pass 1: h' = src + gate * (acc + bias); statistics; h' -> global (TMA store) and back into TMEM ----------
    // The residual tile travels through three 32x32 fp32 staging tiles per warp: TMA load (issued before the
    // accumulator is even ready) -> add in place -> TMA store; the SM never issues a strided global access.
    const int m0 = tile_m * GEMM_BM + q4 * 32;
    const bool warp_valid = m0 < p.M;
    const uint32_t stg = epi_base + (warp - 2) * 12288;
    uint64_t* hb = &hbar[warp - 2][0];
    auto load_h = [&](int c) {   // lane 0 only
      if constexpr (MODE == GLN_RES) {
        const int slot = c % 3;
        mbar_expect_tx(&hb[slot], 4096);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(stg + slot * 4096), "l"(reinterpret_cast<uint64_t>(&tmH)), "r"(smem_u32(&hb[slot])), "r"(col0 + c * 32), "r"(m0)
            : "memory");
      }
    };
    if (lane == 0 && warp_valid) { load_h(0); load_h(1); load_h(2); }
    float4 bq[8], gq[MODE == GLN_RES ? 8 : 1];
    auto prefetch = [&](int col) {
#pragma unroll
      for (int j = 0; j < 8; ++j) bq[j] = __ldg(reinterpret_cast<const float4*>(p.bias + col) + j);
      if constexpr (MODE == GLN_RES) {
#pragma unroll
        for (int j = 0; j < 8; ++j) gq[j] = __ldg(reinterpret_cast<const float4*>(mrow + p.gate_off + col) + j);
      }
    };
    prefetch(col0);
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    float s1 = 0.f, s2 = 0.f;
    if (warp_valid) {
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        const int col = col0 + c * 32, slot = c % 3;
        const uint32_t sb = stg + slot * 4096 + lane * 128;
        uint32_t r[32];
        tmem_ld32(t_addr + c * 32, r);
        if constexpr (MODE == GLN_RES) mbar_wait(&hb[slot], (c / 3) & 1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 hq;
          if constexpr (MODE == GLN_RES) {
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(hq.x), "=f"(hq.y), "=f"(hq.z), "=f"(hq.w)
                         : "r"(sb + ((j ^ (lane & 7)) << 4)));
          } else {
            hq = __ldg(reinterpret_cast<const float4*>(src + col) + j);
          }
          float4 v;
          v.x = __uint_as_float(r[4 * j]) + bq[j].x; v.y = __uint_as_float(r[4 * j + 1]) + bq[j].y;
          v.z = __uint_as_float(r[4 * j + 2]) + bq[j].z; v.w = __uint_as_float(r[4 * j + 3]) + bq[j].w;
          if constexpr (MODE == GLN_RES) {
            v.x = fmaf(gq[j].x, v.x, hq.x); v.y = fmaf(gq[j].y, v.y, hq.y);
            v.z = fmaf(gq[j].z, v.z, hq.z); v.w = fmaf(gq[j].w, v.w, hq.w);
          } else {
            v.x += hq.x; v.y += hq.y; v.z += hq.z; v.w += hq.w;
          }
          s1 += (v.x + v.y) + (v.z + v.w);
          s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
          r[4 * j] = __float_as_uint(v.x); r[4 * j + 1] = __float_as_uint(v.y);
          r[4 * j + 2] = __float_as_uint(v.z); r[4 * j + 3] = __float_as_uint(v.w);
          st_shared_v4(sb + ((j ^ (lane & 7)) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        }
        if (c + 1 < 8) prefetch(col + 32);
        tmem_st32(t_addr + c * 32, r);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmH, stg + slot * 4096, col, m0);
          bulk_commit();
          if constexpr (MODE == GLN_RES) {
            if (c >= 1 && c + 2 < 8) {
              bulk_wait_read<1>();   // the tile stored one chunk ago has been read out: refill it with chunk c + 2
              load_h(c + 2);
            }
          } else {
            bulk_wait_read<2>();     // the tile the next chunk overwrites (stored two chunks ago) has been read out
          }
        }
        __syncwarp();
      }
    }
    tmem_st_wait();
    stats_s[half][r_in] = make_float2(s1, s2);
    named_bar_sync(1, 256);
    const float2 sa_ = stats_s[0][r_in], sb_ = stats_s[1][r_in];
    const float mean = (sa_.x + sb_.x) * (1.0f / GLN_N);
    const float var = fmaxf((sa_.y + sb_.y) * (1.0f / GLN_N) - mean * mean, 0.f);
    const float rstd = rsqrtf(var + 1e-5f);
    if (lane == 0) bulk_wait_read<0>();   // pass 2 reuses the staging tiles
    __syncwarp();

    // ---- pass 2: u = LN(h') * (1 + scale) + shift -> bf16 staging (64 columns = 128 B per row) -> TMA store ----
    const uint32_t stage_hi = stg, stage_lo = stg + 4096;
    float4 cq[8], sq[8];
    auto prefetch2 = [&](int col) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        cq[j] = __ldg(reinterpret_cast<const float4*>(mrow + p.scale_off + col) + j);
        sq[j] = __ldg(reinterpret_cast<const float4*>(mrow + p.shift_off + col) + j);
      }
    };
    prefetch2(col0);
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      const int col = col0 + c * 32;
      const int sub = c & 1;
      uint32_t r[32];
      tmem_ld32(t_addr + c * 32, r);
      tmem_ld_wait();
      float y[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        y[4 * j] = (__uint_as_float(r[4 * j]) - mean) * rstd * (1.f + cq[j].x) + sq[j].x;
        y[4 * j + 1] = (__uint_as_float(r[4 * j + 1]) - mean) * rstd * (1.f + cq[j].y) + sq[j].y;
        y[4 * j + 2] = (__uint_as_float(r[4 * j + 2]) - mean) * rstd * (1.f + cq[j].z) + sq[j].z;
        y[4 * j + 3] = (__uint_as_float(r[4 * j + 3]) - mean) * rstd * (1.f + cq[j].w) + sq[j].w;
      }
      if (c + 1 < 8) prefetch2(col + 32);
      if (sub == 0) {
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
      }
      const uint32_t sbh = stage_hi + lane * 128, sbl = stage_lo + lane * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t h0 = pack_bf16(y[8 * j], y[8 * j + 1]), h1 = pack_bf16(y[8 * j + 2], y[8 * j + 3]);
        const uint32_t h2 = pack_bf16(y[8 * j + 4], y[8 * j + 5]), h3 = pack_bf16(y[8 * j + 6], y[8 * j + 7]);
        const uint32_t off = ((sub * 4 + j) ^ (lane & 7)) << 4;
        st_shared_v4(sbh + off, h0, h1, h2, h3);
        if (p.split3) {
          const uint2 l01 = split_lo4(make_float4(y[8 * j], y[8 * j + 1], y[8 * j + 2], y[8 * j + 3]), make_uint2(h0, h1));
          const uint2 l23 = split_lo4(make_float4(y[8 * j + 4], y[8 * j + 5], y[8 * j + 6], y[8 * j + 7]), make_uint2(h2, h3));
          st_shared_v4(sbl + off, l01.x, l01.y, l23.x, l23.y);
        }
      }
      if (sub == 1) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          const int n0 = col - 32;
          if (m0 < p.M) {
            tma_store_2d(&tmU, stage_hi, n0, m0);
            if (p.split3) {
              tma_store_2d(&tmU, stage_lo, GLN_N + n0, m0);
              tma_store_2d(&tmU, stage_hi, 2 * GLN_N + n0, m0);
            }
          }
          bulk_commit();
        }
      }
    }
    if (lane == 0) bulk_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace stz
