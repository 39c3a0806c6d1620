// gemmln2_kernel — residual-writing GEMM fused with the AdaLN that follows it, as a CLUSTER OF TWO CTAs per
// 128-row block (style denoiser, N = d_model = 512):
//     acc   = A[128, K] · W[512, K]^T            CTA `rank` of the pair computes columns [256 rank, 256 rank + 256)
//     h'    = h + gate[seq] * (acc + b)          (GLN_RES)     or     acc + b + pos[(r/2) % n_style]      (GLN_POS)
//     u     = bf16( LN(h') * (1 + scale[seq]) + shift[seq] )          the next GEMM's A operand ([hi|lo|hi] if split3)
// The LayerNorm needs statistics of full 512-wide rows: each CTA reduces its 256 columns, the two halves of a row
// meet through distributed shared memory (st.shared::cluster into the peer + one cluster barrier), then each CTA
// normalises its own columns straight out of TMEM.  Compared with gemmln_kernel (one CTA per row block, 50 CTAs at
// cfg2) this keeps 100 SMs busy — the same parallelism as the unfused N = 512 GEMM — and removes ln_mod_kernel
// (one launch, one read of h', per sub-layer) and the TMA reduce-add pass.
//
// Warp roles: warp 0 TMA producer (3-stage ring of A 128x64 + W 256x64 tiles), warp 1 MMA issuer (tcgen05.mma
// M 128, N 256), warps 2..9 epilogue: thread = row (TMEM lane), two warps per lane quadrant split the CTA's 256
// columns in halves.  Pass 1: residual tiles TMA-loaded into staging, h' formed in place, TMA-stored, kept in TMEM
// (tcgen05.st), partial sum / sum of squares.  Pass 2: u -> swizzled staging -> TMA store.
#pragma once
#include "gemm_ln.cuh"

namespace stz {

constexpr int GLN2_STAGES = 3, GLN2_BN = 256;
constexpr int GLN2_STAGE_BYTES = GEMM_BM * GEMM_BK * 2 + GLN2_BN * GEMM_BK * 2;   // 48 KB
constexpr int GLN2_SMEM_BYTES = GLN2_STAGES * GLN2_STAGE_BYTES + 8 * 2 * 4096 + 1024;   // + 2 staging tiles per epilogue warp

__device__ __forceinline__ void st_cluster_f32x2(uint32_t local_smem_addr, uint32_t rank, float a, float b) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(rank));
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(a), "f"(b) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(GLN_THREADS, 1) gemmln2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const __grid_constant__ CUtensorMap tmU,
                                                                const __grid_constant__ CUtensorMap tmH,
                                                                const GemmLnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[GLN2_STAGES], empty_bar[GLN2_STAGES], acc_full;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) float2 stats_s[4][GEMM_BM];   // [CTA rank * 2 + column half][row]: (sum, sum of squares)
  __shared__ __align__(8) uint64_t hbar[8][2];           // per epilogue warp: residual chunk landed in staging tile i

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + GLN2_STAGES * GLN2_STAGE_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = p.K / GEMM_BK;
  const int crank = static_cast<int>(g2_cluster_rank());
  const int tile_m = blockIdx.x >> 1;
  const int ncol0 = crank * GLN2_BN;        // first global column of this CTA

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmU);
    prefetch_tmap(&tmH);
    for (int s = 0; s < GLN2_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_full, 1);
    for (int i = 0; i < 16; ++i) mbar_init(&hbar[0][0] + i, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<GLN2_BN>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_sync();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        const uint32_t sa = smem_base + stage * GLN2_STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], GLN2_STAGE_BYTES);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(sa), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(smem_u32(&full_bar[stage])), "r"(kb * GEMM_BK),
            "r"(tile_m * GEMM_BM)
            : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(sa + GEMM_BM * GEMM_BK * 2), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(&full_bar[stage])),
            "r"(kb * GEMM_BK), "r"(ncol0)
            : "memory");
        if (++stage == GLN2_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, GLN2_BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * GLN2_STAGE_BYTES;
        const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + GEMM_BM * GEMM_BK * 2);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == GLN2_STAGES) { stage = 0; phase ^= 1u; }
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
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + half * 128;
    const int col0 = ncol0 + half * 128;            // first global column of this warp

    // ---- pass 1: h' = src + gate * (acc + bias); statistics; h' -> global (TMA store) and back into TMEM ----------
    const int m0 = tile_m * GEMM_BM + q4 * 32;
    const bool warp_valid = m0 < p.M;
    const uint32_t stg = epi_base + (warp - 2) * 8192;
    uint64_t* hb = &hbar[warp - 2][0];
    auto load_h = [&](int c) {   // lane 0 only
      if constexpr (MODE == GLN_RES) {
        const int slot = c & 1;
        mbar_expect_tx(&hb[slot], 4096);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(stg + slot * 4096), "l"(reinterpret_cast<uint64_t>(&tmH)), "r"(smem_u32(&hb[slot])), "r"(col0 + c * 32), "r"(m0)
            : "memory");
      }
    };
    if (lane == 0 && warp_valid) { load_h(0); load_h(1); }
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
      for (int c = 0; c < 4; ++c) {
        const int col = col0 + c * 32, slot = c & 1;
        const uint32_t sb = stg + slot * 4096 + lane * 128;
        uint32_t r[32];
        tmem_ld32(t_addr + c * 32, r);
        if constexpr (MODE == GLN_RES) mbar_wait(&hb[slot], (c >> 1) & 1);
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
        if (c + 1 < 4) prefetch(col + 32);
        tmem_st32(t_addr + c * 32, r);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmH, stg + slot * 4096, col, m0);
          bulk_commit();
          if (c + 2 < 4) {
            bulk_wait_read<0>();     // the tile just stored has been read out: refill it with chunk c + 2 (GLN_RES)
            load_h(c + 2);
          }
        }
        __syncwarp();
      }
    }
    tmem_st_wait();
    // row statistics of this warp's 128 columns -> both CTAs of the pair
    {
      const uint32_t dst = smem_u32(&stats_s[crank * 2 + half][r_in]);
      st_cluster_f32x2(dst, 0, s1, s2);
      st_cluster_f32x2(dst, 1, s1, s2);
    }
  }
  // every thread of both CTAs: the four partial statistics of each row are now visible in both CTAs
  __syncwarp();
  g2_cluster_sync();
  if (warp >= 2) {
    const int q4 = warp & 3, half = (warp - 2) >> 2;
    const int r_in = q4 * 32 + lane;
    const int m = tile_m * GEMM_BM + r_in;
    const int mm = m < p.M ? m : p.M - 1;
    const float* mrow = p.mod + static_cast<size_t>((mm / p.rows_per_utt) * 2 + (mm & 1)) * p.n_mod;
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + half * 128;
    const int col0 = ncol0 + half * 128;
    const int m0 = tile_m * GEMM_BM + q4 * 32;
    const uint32_t stg = epi_base + (warp - 2) * 8192;
    const float2 a0 = stats_s[0][r_in], a1 = stats_s[1][r_in], a2 = stats_s[2][r_in], a3 = stats_s[3][r_in];
    const float mean = ((a0.x + a1.x) + (a2.x + a3.x)) * (1.0f / GLN_N);
    const float var = fmaxf(((a0.y + a1.y) + (a2.y + a3.y)) * (1.0f / GLN_N) - mean * mean, 0.f);
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
    for (int c = 0; c < 4; ++c) {
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
      if (c + 1 < 4) prefetch2(col + 32);
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
  if (warp == 1) tmem_dealloc<GLN2_BN>(tmem_base);
}

}  // namespace stz
