// gemmln3_kernel — a residual-writing GEMM fused with the AdaLN that follows it (style denoiser, N = d_model = 512), as a
// CLUSTER OF TWO CTAs per 128-row block:
//     acc   = A[128, K] · W[512, K]^T            CTA `rank` of the pair computes columns [256 rank, 256 rank + 256)
//     h'    = h + gate[seq] * (acc + b)          (GLN_RES)     or     acc + b + pos[(r/2) % n_style]      (GLN_POS)
//     u     = bf16( LN(h') * (1 + scale[seq]) + shift[seq] )          the next GEMM's A operand ([hi|lo|hi] if split3)
// This replaces {GEMM with gated-residual epilogue, ln_mod_kernel}: one launch instead of two, h' written once, u produced
// without re-reading h' from global memory.  The LayerNorm needs statistics of full 512-wide rows: each CTA reduces its
// 256 columns, the halves of a row meet through distributed shared memory (st.shared::cluster + one cluster barrier).
// Two earlier forms (one CTA per row block; cluster of two with per-chunk residual staging) measured slower and were
// removed (git history: gemm_ln.cuh, gemm_ln2.cuh).  What this version does differently:
//   * once the producer has issued the last operand tiles it keeps going around the ring and TMA-loads the CTA's whole
//     residual tile h[128 x 256] fp32 (8 boxes of 128 rows x 32 columns = 128 KB) into the stages as the MMAs release
//     them — under the tail of the mainloop, so the tile is resident when the accumulator is;
//   * pass 1 forms h' = h + gate * (acc + b) IN PLACE in that tile, accumulates the row statistics, and four TMA stores
//     per column half write h' back — no per-chunk waits;
//   * pass 2 re-reads h' (kept in TMEM), normalises + modulates, and builds the bf16 operand u in the remaining 64 KB of
//     the ring (4 boxes of 128 rows x 64 columns), stored with two TMA stores per column half.
//   * bias and the per-sequence gate / scale / shift rows of the tile (<= 8 sequences x 256 columns) are copied into a
//     25 KB shared-memory table by the epilogue warps WHILE the mainloop runs: fetched from L2 inside the passes they
//     cost a full ~1.5 k-cycle round trip per 32-column chunk (measured: 13 k + 9 k cycles for the two passes).
// Shared memory = the 4-stage operand ring (192 KB) + that table.
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
  int single;             // row layout (see GemmParams::single)
  int tile_rows;          // rows of the residual stream one CTA pair owns: 128, or fewer (a multiple of 8) so that a problem
                          // with 38 .. 74 full tiles spreads over (nearly) all 148 SMs — the MMAs still run at M = 128 (the tile's
                          // unused accumulator rows are never read), the shared-memory-bound epilogue passes shrink with the rows
};

constexpr int GLN_N = 512, GLN_THREADS = 320;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void st_cluster_f32x2(uint32_t local_smem_addr, uint32_t rank, float a, float b) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(rank));
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(a), "f"(b) : "memory");
}

constexpr int GLN3_STAGES = 4, GLN3_BN = 256;
constexpr int GLN3_STAGE_BYTES = GEMM_BM * GEMM_BK * 2 + GLN3_BN * GEMM_BK * 2;   // 48 KB
constexpr int GLN3_RING_BYTES = GLN3_STAGES * GLN3_STAGE_BYTES;                   // 192 KB
constexpr int GLN3_H_BOX = 128 * 128;                                              // 128 rows x 32 fp32 (or 64 bf16)
constexpr int GLN3_H_BYTES = 8 * GLN3_H_BOX;                                       // residual tile: 128 KB
constexpr int GLN3_SMEM_BYTES = GLN3_RING_BYTES + 1024;
constexpr int GLN3_MAX_SEQ = 8;       // sequences (utterance x CFG branch) a 128-row tile may span: rows_per_utt >= 64
static_assert(GLN3_H_BYTES + 4 * GLN3_H_BOX == GLN3_RING_BYTES, "residual tile + u tile fill the ring exactly");

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

template <int MODE>
__global__ void __launch_bounds__(GLN_THREADS, 1) gemmln3_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const __grid_constant__ CUtensorMap tmU,
                                                                const __grid_constant__ CUtensorMap tmH,
                                                                const GemmLnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[GLN3_STAGES], empty_bar[GLN3_STAGES], acc_full, h_full;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) float2 stats_s[4][GEMM_BM];   // [CTA rank * 2 + column half][row]: (sum, sum of squares)
  __shared__ __align__(16) float tab_s[3][GLN3_MAX_SEQ][GLN3_BN];   // gate | scale | shift rows of the tile's sequences
  __shared__ __align__(16) float bias_s[GLN3_BN];

  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t u_base = ring + GLN3_H_BYTES;
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp index provably uniform
#ifdef STZ_TRACE
  long long* tr = (g_gemm_trace != nullptr && warp == 2 && lane == 0) ? g_gemm_trace + (148 + blockIdx.x) * 64 : nullptr;   // second half of the buffer (gemm2_kernel uses the first)
#else
  constexpr long long* tr = nullptr;
#endif
  if (tr != nullptr) tr[0] = clock64();
  const int num_kb = p.K / GEMM_BK;                 // multiple of GLN3_STAGES (host-checked)
  const int TR = p.tile_rows;                       // valid rows of this tile (host: 8 | TR, 8 <= TR <= 128)
  const uint32_t a_bytes = static_cast<uint32_t>(TR) * 128u;                       // one A box: TR rows x 64 bf16
  const uint32_t stage_tx = a_bytes + GLN3_BN * GEMM_BK * 2, h_tx = 8u * a_bytes;   // (an h box is TR rows x 32 fp32: the same bytes)
  const int crank = static_cast<int>(g2_cluster_rank());
  const int tile_m = blockIdx.x >> 1;
  const int ncol0 = crank * GLN3_BN;                // first global column of this CTA
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmU);
    prefetch_tmap(&tmH);
    for (int s = 0; s < GLN3_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&h_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<GLN3_BN>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // the first ring of (constant) W tiles goes out before the dependency wait, as in gemm2_kernel
  int n_pre = 0;
  if (warp == 0 && lane == 0) {
    n_pre = num_kb < GLN3_STAGES ? num_kb : GLN3_STAGES;
    for (int s = 0; s < n_pre; ++s) {
      mbar_expect_tx(&full_bar[s], stage_tx);
      tma_load_2d_u32(ring + s * GLN3_STAGE_BYTES + A_BYTES, &tmB, smem_u32(&full_bar[s]), s * GEMM_BK, ncol0);
    }
  }
  pdl_sync();
  if (tr != nullptr) tr[1] = clock64();

  if (warp == 0) {
    if (lane == 0) {     // TMA producer: one lane (see gemm2_kernel)
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (kb >= GLN3_STAGES) mbar_wait(&empty_bar[stage], phase ^ 1u);   // (the first ring is free by construction: a passing
                                                                           // try_wait still costs the issuing lane ~200 cycles)
        const uint32_t sa = ring + stage * GLN3_STAGE_BYTES;
        if (kb >= n_pre) {
          mbar_expect_tx(&full_bar[stage], stage_tx);
          tma_load_2d_u32(sa + A_BYTES, &tmB, smem_u32(&full_bar[stage]), kb * GEMM_BK, ncol0);
        }
        tma_load_2d_u32(sa, &tmA, smem_u32(&full_bar[stage]), kb * GEMM_BK, tile_m * TR);
        if (++stage == GLN3_STAGES) { stage = 0; phase ^= 1u; }
      }
      if constexpr (MODE == GLN_RES) {
        // residual tile into the ring as the MMAs release it: boxes 0..2 live in stage 0, 3..5 in stage 1, 6..7 in stage 2
        mbar_expect_tx(&h_full, h_tx);
        for (int s = 0; s < 3; ++s) {            // stage == s here: num_kb is a multiple of the ring depth
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          for (int cc = 3 * s; cc < 3 * s + 3 && cc < 8; ++cc)
            tma_load_2d_u32(ring + cc * GLN3_H_BOX, &tmH, smem_u32(&h_full), ncol0 + cc * 32, tile_m * TR);
          if (++stage == GLN3_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform loop, one elected lane issues (see gemm2_kernel)
    constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, GLN3_BN);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sa = ring + stage * GLN3_STAGE_BYTES;
      const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + A_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == GLN3_STAGES) { stage = 0; phase ^= 1u; }
    }
    if (elect_one()) umma_commit(&acc_full);
    __syncwarp();
  } else {
    const int q4 = warp & 3, half = (warp - 2) >> 2;
    const int r_in = q4 * 32 + lane;                // row inside the tile == TMEM lane
    const int m = tile_m * TR + r_in;
    const int mm = m < p.M ? m : p.M - 1;           // clamp for loads; rows >= M (or >= TR inside the tile) are never stored (TMA clips)
    const bool warp_active = q4 * 32 < TR;          // warps whose 32 rows lie past the tile's rows skip both passes
    const float* src = p.pos + static_cast<size_t>(tok_of_row(mm, p.n_style, p.single)) * GLN_N;   // GLN_POS only
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + half * 128;
    const int col0 = ncol0 + half * 128;            // first global column of this warp
    const uint32_t sw = r_in & 7;
    // modulation table: every epilogue thread copies its share while the mainloop runs
    const int m_first = tile_m * TR < p.M ? tile_m * TR : p.M - 1;
    const int seq_first = seq_of_row(m_first & ~1, p.rows_per_utt, p.single), seq_last = seq_of_row((p.M - 1) | (p.single ? 0 : 1), p.rows_per_utt, p.single);
    {
      const int te = threadIdx.x - 64;                     // 0..255
      for (int idx = te; idx < 3 * GLN3_MAX_SEQ * (GLN3_BN / 4); idx += 256) {
        const int kind = idx / (GLN3_MAX_SEQ * (GLN3_BN / 4)), sl = (idx / (GLN3_BN / 4)) % GLN3_MAX_SEQ, c4 = idx % (GLN3_BN / 4);
        if (MODE != GLN_RES && kind == 0) continue;
        const int seq = seq_first + sl < seq_last ? seq_first + sl : seq_last;
        const int off = kind == 0 ? p.gate_off : (kind == 1 ? p.scale_off : p.shift_off);
        reinterpret_cast<float4*>(&tab_s[kind][sl][0])[c4] =
            __ldg(reinterpret_cast<const float4*>(p.mod + static_cast<size_t>(seq) * p.n_mod + off + ncol0) + c4);
      }
      if (te < GLN3_BN / 4) reinterpret_cast<float4*>(bias_s)[te] = __ldg(reinterpret_cast<const float4*>(p.bias + ncol0) + te);
      named_bar_sync(1, 256);
    }
    const int sl = seq_of_row(mm, p.rows_per_utt, p.single) - seq_first;      // this row's sequence inside the table
    const uint32_t gate_t = smem_u32(&tab_s[0][sl][half * 128]), bias_t = smem_u32(&bias_s[half * 128]);
    mbar_wait(&acc_full, 0);
    if (tr != nullptr) tr[2] = clock64();
    if constexpr (MODE == GLN_RES) mbar_wait(&h_full, 0);
    if (tr != nullptr) tr[3] = clock64();
    tc_fence_after();
    // ---- pass 1: h' = h + gate * (acc + bias)  (or acc + bias + pos), in place in the ring tile; statistics
    float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
    for (int c = 0; c < (warp_active ? 4 : 0); ++c) {
      const int col = col0 + c * 32;
      const uint32_t sb = ring + (half * 4 + c) * GLN3_H_BOX + r_in * 128;
      uint32_t r[32];
      if (tr != nullptr && c < 2) tr[16 + 4 * c] = clock64();
      tmem_ld32(t_addr + c * 32, r);
      tmem_ld_wait();
      if (tr != nullptr && c < 2) tr[17 + 4 * c] = clock64();
      // all shared-memory reads of the chunk first (independent, they pipeline), then the arithmetic, then the stores:
      // interleaved per 16-byte piece the chunk was a chain of 8 x (load latency + math + store)
      float4 hq[8], bqv[8], gqv[MODE == GLN_RES ? 8 : 1];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if constexpr (MODE == GLN_RES) hq[j] = lds_f4(sb + ((j ^ sw) << 4));
        else hq[j] = __ldg(reinterpret_cast<const float4*>(src + col) + j);
        bqv[j] = lds_f4(bias_t + (c * 32 + j * 4) * 4);
        if constexpr (MODE == GLN_RES) gqv[j] = lds_f4(gate_t + (c * 32 + j * 4) * 4);
      }
      float4 vv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 v;
        v.x = __uint_as_float(r[4 * j]) + bqv[j].x; v.y = __uint_as_float(r[4 * j + 1]) + bqv[j].y;
        v.z = __uint_as_float(r[4 * j + 2]) + bqv[j].z; v.w = __uint_as_float(r[4 * j + 3]) + bqv[j].w;
        if constexpr (MODE == GLN_RES) {
          v.x = fmaf(gqv[j].x, v.x, hq[j].x); v.y = fmaf(gqv[j].y, v.y, hq[j].y);
          v.z = fmaf(gqv[j].z, v.z, hq[j].z); v.w = fmaf(gqv[j].w, v.w, hq[j].w);
        } else {
          v.x += hq[j].x; v.y += hq[j].y; v.z += hq[j].z; v.w += hq[j].w;
        }
        s1 += (v.x + v.y) + (v.z + v.w);
        s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        vv[j] = v;
        r[4 * j] = __float_as_uint(v.x); r[4 * j + 1] = __float_as_uint(v.y);
        r[4 * j + 2] = __float_as_uint(v.z); r[4 * j + 3] = __float_as_uint(v.w);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        st_shared_v4(sb + ((j ^ sw) << 4), __float_as_uint(vv[j].x), __float_as_uint(vv[j].y), __float_as_uint(vv[j].z), __float_as_uint(vv[j].w));
      tmem_st32(t_addr + c * 32, r);      // h' also replaces the accumulator: pass 2 re-reads it from TMEM, not from the
                                          // shared-memory tile the TMA store engine is draining
      if (tr != nullptr && c < 2) tr[18 + 4 * c] = clock64();
    }
    if (tr != nullptr) tr[4] = clock64();
    {  // row statistics of this warp's 128 columns -> both CTAs of the pair
      const uint32_t dst = smem_u32(&stats_s[crank * 2 + half][r_in]);
      st_cluster_f32x2(dst, 0, s1, s2);
      st_cluster_f32x2(dst, 1, s1, s2);
    }
    tmem_st_wait();
    fence_proxy_async();                       // h' was written through the generic proxy, TMA reads it through the async proxy
    if (tr != nullptr) tr[5] = clock64();
  }
  // every thread of both CTAs: the four partial statistics of each row are now visible in both CTAs
  __syncwarp();
  g2_cluster_sync();
  if (tr != nullptr) tr[6] = clock64();
  if (warp >= 2) {
    const int q4 = warp & 3, half = (warp - 2) >> 2;
    const int r_in = q4 * 32 + lane;
    const int m = tile_m * TR + r_in;
    const int mm = m < p.M ? m : p.M - 1;
    const bool warp_active = q4 * 32 < TR;
    const int col0 = ncol0 + half * 128;
    const uint32_t sw = r_in & 7;
    const float2 a0 = stats_s[0][r_in], a1 = stats_s[1][r_in], a2 = stats_s[2][r_in], a3 = stats_s[3][r_in];
    const float mean = ((a0.x + a1.x) + (a2.x + a3.x)) * (1.0f / GLN_N);
    const float var = fmaxf(((a0.y + a1.y) + (a2.y + a3.y)) * (1.0f / GLN_N) - mean * mean, 0.f);
    const float rstd = rsqrtf(var + 1e-5f);
    // h' goes out now (the cluster barrier above also ordered the four row groups of each column half; issued before
    // it, the barrier's release waited ~5 k cycles for the 128 KB of stores to drain)
    if (q4 == 0 && elect_one()) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tma_store_2d(&tmH, ring + (half * 4 + c) * GLN3_H_BOX, col0 + c * 32, tile_m * TR);
      bulk_commit();
    }
    if (p.split3) {     // the lo tiles reuse the residual boxes: their TMA stores must have read them out first
      if (q4 == 0 && elect_one()) bulk_wait_read<0>();
      named_bar_sync(2 + half, 128);
    }
    // ---- pass 2: u = LN(h') * (1 + scale) + shift -> bf16 tile (128 rows x 64 columns per box) in the rest of the ring
    const int m_first = tile_m * TR < p.M ? tile_m * TR : p.M - 1;
    const int sl = seq_of_row(mm, p.rows_per_utt, p.single) - seq_of_row(m_first & ~1, p.rows_per_utt, p.single);
    const uint32_t scale_t = smem_u32(&tab_s[1][sl][half * 128]), shift_t = smem_u32(&tab_s[2][sl][half * 128]);
    const uint32_t t_addr2 = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + half * 128;
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < (warp_active ? 4 : 0); ++c) {
      const int sub = c & 1, ub = half * 2 + (c >> 1);
      uint32_t hr[32];
      tmem_ld32(t_addr2 + c * 32, hr);
      float4 cqv[8], sqv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        cqv[j] = lds_f4(scale_t + (c * 32 + j * 4) * 4);
        sqv[j] = lds_f4(shift_t + (c * 32 + j * 4) * 4);
      }
      tmem_ld_wait();
      float y[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        y[4 * j] = (__uint_as_float(hr[4 * j]) - mean) * rstd * (1.f + cqv[j].x) + sqv[j].x;
        y[4 * j + 1] = (__uint_as_float(hr[4 * j + 1]) - mean) * rstd * (1.f + cqv[j].y) + sqv[j].y;
        y[4 * j + 2] = (__uint_as_float(hr[4 * j + 2]) - mean) * rstd * (1.f + cqv[j].z) + sqv[j].z;
        y[4 * j + 3] = (__uint_as_float(hr[4 * j + 3]) - mean) * rstd * (1.f + cqv[j].w) + sqv[j].w;
      }
      const uint32_t ubh = u_base + ub * GLN3_H_BOX + r_in * 128;
      const uint32_t ubl = ring + (2 * ub) * GLN3_H_BOX + r_in * 128;     // split3: lo tile over residual box 2 ub (see above)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t h0 = pack_bf16(y[8 * j], y[8 * j + 1]), h1 = pack_bf16(y[8 * j + 2], y[8 * j + 3]);
        const uint32_t h2 = pack_bf16(y[8 * j + 4], y[8 * j + 5]), h3 = pack_bf16(y[8 * j + 6], y[8 * j + 7]);
        const uint32_t off = ((sub * 4 + j) ^ sw) << 4;
        st_shared_v4(ubh + off, h0, h1, h2, h3);
        if (p.split3) {
          const uint2 l01 = split_lo4(make_float4(y[8 * j], y[8 * j + 1], y[8 * j + 2], y[8 * j + 3]), make_uint2(h0, h1));
          const uint2 l23 = split_lo4(make_float4(y[8 * j + 4], y[8 * j + 5], y[8 * j + 6], y[8 * j + 7]), make_uint2(h2, h3));
          st_shared_v4(ubl + off, l01.x, l01.y, l23.x, l23.y);
        }
      }
    }
    if (tr != nullptr) tr[7] = clock64();
    fence_proxy_async();
    named_bar_sync(2 + half, 128);
    if (tr != nullptr) tr[8] = clock64();
    if (q4 == 0 && elect_one()) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int ub = half * 2 + k, n0 = ncol0 + ub * 64;
        tma_store_2d(&tmU, u_base + ub * GLN3_H_BOX, n0, tile_m * TR);
        if (p.split3) {
          tma_store_2d(&tmU, ring + (2 * ub) * GLN3_H_BOX, GLN_N + n0, tile_m * TR);
          tma_store_2d(&tmU, u_base + ub * GLN3_H_BOX, 2 * GLN_N + n0, tile_m * TR);
        }
      }
      bulk_commit();
      bulk_wait_read<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tr != nullptr) tr[9] = clock64();
  if (warp == 1) tmem_dealloc<GLN3_BN>(tmem_base);
}

}  // namespace stz
