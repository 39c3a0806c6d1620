// gemm2_kernel — the product GEMM of the style denoiser: C[M,N] = A[M,K] · W[N,K]^T (+ fused epilogue).
//
// Persistent, warp-specialised, one CTA per SM (grid = min(#tiles, #SMs)):
//   warp 0      TMA producer: 128B-swizzled K-major tiles of A (128 x 64) and W (BN x 64) into a smem ring
//   warp 1      TMEM owner + MMA issuer: tcgen05.mma kind::f16 (bf16 x bf16 -> fp32), 128 x BN accumulator,
//               DOUBLE-BUFFERED in TMEM (2 x BN columns) so the epilogue of tile i overlaps the mainloop of i+1
//   warps 2..9  epilogue (two per TMEM lane quadrant, alternating column chunks): tcgen05.ld (thread = row) -> bias / GELU / gate -> 128B-swizzled smem staging ->
//               TMA store (cp.async.bulk.tensor) or, for the gated residual update h += gate * (acc + b),
//               TMA reduce-add (cp.reduce.async.bulk.tensor .add.f32): h is never read by an SM.
// Barriers: full/empty per smem stage (TMA <-> MMA), tmem_full/tmem_empty per accumulator (MMA <-> epilogue).
//
// Tiles are visited n-fastest so CTAs running concurrently share A row-blocks in L2.  Rows past M are
// zero-filled on load and clipped on store by the tensor maps (no tail code).
#pragma once
#include <cuda_fp16.h>
#include "gemm.cuh"

namespace stz {

constexpr int G2_THREADS = 320;   // TMA warp + MMA warp + 8 epilogue warps
constexpr int G2_STAGE_BYTES_EPI = 8 * 4096;  // 8 epilogue warps x one staging tile (32 rows x 128 B)

template <int BN>
constexpr int g2_stages() { return BN == 256 ? 4 : (BN == 192 ? 4 : 6); }
template <int BN>
constexpr int g2_smem_bytes() { return g2_stages<BN>() * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2) + G2_STAGE_BYTES_EPI + 1024; }
template <int BN>
constexpr int g2_tmem_cols() { return BN == 128 ? 256 : 512; }

// ---- TMA store / reduce helpers -----------------------------------------------------------------
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const void* tmap, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.0f + tanh_fast(k0 * (x + k1 * x * x * x)));
}

// GELU-tanh of two values at once in packed fp16 arithmetic -> packed bf16.  The FFN1 epilogue is instruction-issue
// bound (tools/gemm_trace.py: a 128 x 256 tile's epilogue took longer than its mainloop and slowed it by ~45 %);
// fp16x2 halves the arithmetic instructions.  fp16 carries 11 significant bits against the 8 of the bf16 result,
// |x| of the FFN pre-activations is far inside the fp16 range, and tiny |x| only lose absolute accuracy (< 6e-8).
__device__ __forceinline__ uint32_t gelu_tanh_f16x2_to_bf16x2(float a, float b) {
  const __half2 x = __floats2half2_rn(a, b);
  const __half2 x2 = __hmul2(x, x);
  const __half2 inner = __hmul2(x, __hfma2(x2, __float2half2_rn(0.7978845608028654f * 0.044715f), __float2half2_rn(0.7978845608028654f)));
  __half2 t;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(*reinterpret_cast<uint32_t*>(&t)) : "r"(*reinterpret_cast<const uint32_t*>(&inner)));
  const __half2 hx = __hmul2(x, __float2half2_rn(0.5f));
  const __half2 y = __hfma2(hx, t, hx);
  const float2 f = __half22float2(y);
  return pack_bf16(f.x, f.y);
}

// Staged epilogues (TMA store / reduce): EPI_F32, EPI_F32_POS, EPI_BF16, EPI_GELU_BF16, EPI_GATE_RES.
// EPI_SAMPLER: sampler_epilogue32 below (warp-transposed through the staging tile, coalesced global access).

// CFG combine + EDM/sampler affine update + next input for one warp: tile rows m0 .. m0 + 31 (lane = row, R layout:
// even rows conditional, odd rows unconditional) x 32 columns from n0, accumulators (+ bias) in v[].
// The guided F of the 16 state rows is transposed through a 2 KB shared-memory tile so that the state update runs
// with lane = (state row, 16-column half): every global access is a full 32-byte sector run (the row-per-thread
// form issued 24 scattered 8-byte stores per thread and chunk and was bound by L1 store transactions).
// Single-branch layout (guidance-conditioned student): no CFG pair — every tile row is a state row; the warp's 32 rows go
// through the same transposed update in two halves of 16.
__device__ __forceinline__ void sampler_epilogue32(const GemmParams& p, int m0, int n0, float (&v)[32], uint32_t stage, int lane) {
  if (!p.single) {
    const float w = __ldg(p.coef + 5);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float other = __shfl_xor_sync(0xffffffffu, v[j], 1);
      const float Fc = (lane & 1) ? other : v[j], Fu = (lane & 1) ? v[j] : other;
      v[j] = Fu + w * (Fc - Fu);
    }
  }
  const int n_half = p.single ? 2 : 1;
  for (int hf = 0; hf < n_half; ++hf) {
  if (p.single ? (lane >> 4) == hf : (lane & 1) == 0) {
    const int r = p.single ? (lane & 15) : (lane >> 1);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      st_shared_v4(stage + r * 128 + ((j ^ (r & 7)) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                   __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
  }
  __syncwarp();
  const int r = lane >> 1, hs = lane & 1;
  const int srow = p.single ? m0 + 16 * hf + r : (m0 >> 1) + r;                       // state row
  const int nrep = p.single ? 1 : 2;                    // activation rows fed by one state row
  if (nrep * srow < p.M) {
    const float cx = __ldg(p.coef + 0), cm = __ldg(p.coef + 1), cF = __ldg(p.coef + 2), cn = __ldg(p.coef + 3);
    const float cin = __ldg(p.coef + 4);
    const bool to_mid = __ldg(p.coef + 6) != 0.0f;
    const size_t so = static_cast<size_t>(srow) * p.N + n0 + hs * 16;
    float* dst = (to_mid ? p.xmid : p.x) + so;
    __nv_bfloat16* xo = p.xin + static_cast<size_t>(nrep * srow) * 3 * p.N + n0 + hs * 16;
#pragma unroll
    for (int jj = 0; jj < 4; jj += 2) {
      float4 o[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float4 F;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(F.x), "=f"(F.y), "=f"(F.z), "=f"(F.w)
                     : "r"(stage + r * 128 + (((4 * hs + jj + q) ^ (r & 7)) << 4)));
        const float4 xv = *reinterpret_cast<const float4*>(p.x + so + (jj + q) * 4);
        o[q] = make_float4(cx * xv.x + cF * F.x, cx * xv.y + cF * F.y, cx * xv.z + cF * F.z, cx * xv.w + cF * F.w);
        if (cm != 0.0f) {
          const float4 xm = *reinterpret_cast<const float4*>(p.xmid + so + (jj + q) * 4);
          o[q].x += cm * xm.x; o[q].y += cm * xm.y; o[q].z += cm * xm.z; o[q].w += cm * xm.w;
        }
        if (cn != 0.0f) {
          const float4 nz = __ldg(reinterpret_cast<const float4*>(p.noise + so + (jj + q) * 4));
          o[q].x += cn * nz.x; o[q].y += cn * nz.y; o[q].z += cn * nz.z; o[q].w += cn * nz.w;
        }
        if (p.tap != nullptr) *reinterpret_cast<float4*>(p.tap + so + (jj + q) * 4) = F;
      }
      *reinterpret_cast<float4*>(dst + jj * 4) = o[0];
      *reinterpret_cast<float4*>(dst + jj * 4 + 4) = o[1];
      // next eval's input c_in(sigma') x' as the split-bf16 A operand [hi | lo | hi] (row stride 3N), both CFG branches
      const float4 y0 = make_float4(cin * o[0].x, cin * o[0].y, cin * o[0].z, cin * o[0].w);
      const float4 y1 = make_float4(cin * o[1].x, cin * o[1].y, cin * o[1].z, cin * o[1].w);
      uint2 h0, h1;
      h0.x = pack_bf16(y0.x, y0.y); h0.y = pack_bf16(y0.z, y0.w);
      h1.x = pack_bf16(y1.x, y1.y); h1.y = pack_bf16(y1.z, y1.w);
      const uint2 l0 = split_lo4(y0, h0), l1 = split_lo4(y1, h1);
      const uint4 hi = make_uint4(h0.x, h0.y, h1.x, h1.y), lo = make_uint4(l0.x, l0.y, l1.x, l1.y);
      for (int b = 0; b < nrep; ++b) {
        __nv_bfloat16* xr = xo + static_cast<size_t>(b) * 3 * p.N + jj * 4;
        *reinterpret_cast<uint4*>(xr) = hi;
        *reinterpret_cast<uint4*>(xr + p.N) = lo;
        *reinterpret_cast<uint4*>(xr + 2 * p.N) = hi;
      }
    }
  }
  __syncwarp();   // the staging tile is reused by this warp's next half / chunk
  }
}
template <int EPI>
constexpr bool g2_staged() { return EPI == EPI_F32 || EPI == EPI_F32_POS || EPI == EPI_BF16 || EPI == EPI_GELU_BF16 || EPI == EPI_GATE_RES; }
template <int EPI>
constexpr bool g2_out_bf16() { return EPI == EPI_BF16 || EPI == EPI_GELU_BF16; }

// CM = 2 — CTA pair (tcgen05 cta_group::2): a cluster of two CTAs computes one 256 x BN tile.  Each CTA stages its own
// 128 rows of A and HALF of the W tile (BN/2 rows), so shared-memory traffic per SM per MMA drops from
// (128 + BN) to (128 + BN/2) operand rows — the single-CTA kernel is bound by shared-memory bandwidth (operand reads
// by the tensor core + TMA writes + epilogue staging > 128 B/clk).  The leader CTA (rank 0) issues
// tcgen05.mma.cta_group::2 (M = 256) for the pair; both CTAs' TMA loads signal the leader's full barrier; stages and
// accumulators are handed back with multicast tcgen05.commit; each CTA drains its own 128 TMEM lanes.
__device__ __forceinline__ uint32_t g2_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void g2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t g2_mapa(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(rank));
  return remote;
}
__device__ __forceinline__ void g2_remote_arrive(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs of this thread completed) on the same barrier offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}
// TMA load into this CTA's shared memory whose completion bytes are posted on a barrier of the pair's leader
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t smem_dst, const void* tmap, uint32_t cluster_bar_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(cluster_bar_addr), "r"(c0), "r"(c1)
      : "memory");
}

template <int BN, int CM>
constexpr int g2_stage_bytes() { return GEMM_BM * GEMM_BK * 2 + (BN / CM) * GEMM_BK * 2; }
template <int BN, int CM>
constexpr int g2_stages_cm() { return CM == 1 ? g2_stages<BN>() : (BN == 128 ? 8 : 6); }
template <int BN, int CM>
constexpr int g2_smem_bytes_cm() { return g2_stages_cm<BN, CM>() * g2_stage_bytes<BN, CM>() + G2_STAGE_BYTES_EPI + 1024; }

// Debug timeline (tools/gemm_trace.py): when set, every CTA records clock64() stamps into g_gemm_trace[cta][64]:
// [0] start, [1] after TMEM alloc / barrier init, [2] after griddepcontrol.wait, [3] producer: first TMA issued,
// [4] producer: last TMA issued; per tile t < 4: [8+4t] MMA: first stage landed, [9+4t] MMA: tile committed,
// [10+4t] epilogue warp 2: accumulator ready, [11+4t] epilogue warp 2: tile drained; [5] end.
__device__ long long* g_gemm_trace = nullptr;
__device__ int g_gemm_dbg = 0;   // experiment flags (tools/gemm_trace.py): 1 skip TMA store, 2 skip GELU, 4 skip staging stores

template <int BN, int EPI, int CM = 1>
__global__ void __launch_bounds__(G2_THREADS, 1) gemm2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB,
                                                              const __grid_constant__ CUtensorMap tmC,
                                                              const GemmParams p) {
  constexpr int STAGES = g2_stages_cm<BN, CM>();
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  constexpr int B_BYTES = (BN / CM) * GEMM_BK * 2;        // this CTA's share of the W tile
  constexpr int TMEM_COLS = g2_tmem_cols<BN>();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full[2];
  __shared__ __align__(8) uint64_t tmem_empty[2];
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + STAGES * (A_BYTES + B_BYTES);
  // the warp index through a shuffle: provably warp-uniform for the compiler, so everything derived from it (staging-tile
  // addresses, store coordinates) lives in uniform registers and the epilogue's TMA stores need no ELECT / R2UR waterfall
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
#ifdef STZ_TRACE
  long long* tr = g_gemm_trace != nullptr ? g_gemm_trace + blockIdx.x * 64 : nullptr;
  const int dbg = g_gemm_dbg;
#else
  constexpr long long* tr = nullptr;   // timeline / experiment hooks compile away (make TRACE=1 enables them)
  constexpr int dbg = 0;
#endif
  if (tr != nullptr && threadIdx.x == 0) tr[0] = clock64();
  const int num_kb = p.K / GEMM_BK;
  const int tiles_n = p.N / BN;
  // work units: (CM * 128) rows x BN columns; unit u -> rows of this CTA: tile_m = CM * (u / tiles_n) + rank
  const int n_tiles = tiles_n * (((p.M + GEMM_BM - 1) / GEMM_BM + CM - 1) / CM);
  const int crank = CM > 1 ? static_cast<int>(g2_cluster_rank()) : 0;
  const bool leader = crank == 0;
  const int unit0 = blockIdx.x / CM, unit_stride = gridDim.x / CM;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (g2_staged<EPI>()) prefetch_tmap(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);           // pair: the leader arms the bytes of BOTH CTAs; the peer's TMA posts its bytes remotely
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 8 * CM);    // 8 epilogue warps per CTA
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CM == 1) tmem_alloc<TMEM_COLS>(&tmem_slot);
    else tmem_alloc2<TMEM_COLS>(&tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if constexpr (CM > 1) g2_cluster_sync();   // peer barriers are initialised before any remote arrive / commit
  if (tr != nullptr && threadIdx.x == 0) tr[1] = clock64();
  // The W operand is constant (weights): the first ring of W tiles is requested BEFORE waiting for the previous kernel,
  // so only the A tiles (its output) are fetched on the critical path after the dependency resolves.
  constexpr int kPre = STAGES;
  int n_pre = 0;
  // first tile's coordinates: the integer division runs here, under the previous kernel's tail, not behind the dependency wait
  const int first_q = unit0 / tiles_n, first_r = unit0 - first_q * tiles_n;
  if constexpr (CM == 1) {
    if (warp == 0 && lane == 0 && unit0 < n_tiles && p.tile_needed == nullptr) {   // (the skip list is produced upstream: not readable yet)
      n_pre = num_kb < kPre ? num_kb : kPre;
      const int tile_n0 = first_r;
      for (int s = 0; s < n_pre; ++s) {
        mbar_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(smem_base + s * (A_BYTES + B_BYTES) + A_BYTES), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(&full_bar[s])),
            "r"(s * GEMM_BK), "r"(tile_n0 * BN)
            : "memory");
      }
    }
  }
  pdl_sync();   // everything above overlapped the previous kernel's tail; activations are touched only below
  if (tr != nullptr && threadIdx.x == 0) tr[2] = clock64();

  if (warp == 0) {
    // TMA producer: one lane.  (The warp-uniform + elect form that pays for the MMA warp below was measured SLOWER here:
    // 688 vs 568 cycles per k-block, tools/mainloop_probe.py — two TMA issues per k-block do not pace the loop.)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ring_fresh = true;     // no stage has been handed to the MMA warp yet: the first ring needs no empty-barrier wait
      for (int tile = unit0; tile < n_tiles; tile += unit_stride) {
        const int tq = tile == unit0 ? first_q : tile / tiles_n;
        const int tile_m = CM * tq + crank, tile_n = tile == unit0 ? first_r : tile - tq * tiles_n;
        if (CM == 1 && p.tile_needed != nullptr && p.tile_needed[tile_m] == 0) continue;   // all-padding rows
        for (int kb = 0; kb < num_kb; ++kb) {
          // (a passing mbarrier.try_wait still costs the issuing lane ~200 cycles: skipped while the ring is fresh)
          if (!(ring_fresh && kb < STAGES)) mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (kb == num_kb - 1) ring_fresh = false;
          const uint32_t sa = smem_base + stage * (A_BYTES + B_BYTES);
          if constexpr (CM == 1) {
            const bool pre = tile == unit0 && kb < n_pre;   // this stage's W tile (and its expect_tx) went out before the wait
            if ((dbg & 24) && !pre) {   // timing experiments (trace build): 8 = no W loads, 16 = no A loads (results are wrong)
              const uint32_t bytes = ((dbg & 8) ? 0u : B_BYTES) + ((dbg & 16) ? 0u : A_BYTES);
              if (bytes == 0) { mbar_arrive(&full_bar[stage]); }
              else {
                mbar_expect_tx(&full_bar[stage], bytes);
                if (!(dbg & 16))
                  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                               ::"r"(sa), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(smem_u32(&full_bar[stage])), "r"(kb * GEMM_BK), "r"(p.a_row0 + tile_m * GEMM_BM) : "memory");
                if (!(dbg & 8))
                  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                               ::"r"(sa + A_BYTES), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(&full_bar[stage])), "r"(kb * GEMM_BK), "r"(tile_n * BN) : "memory");
              }
              if (++stage == STAGES) { stage = 0; phase ^= 1u; }
              continue;
            }
            if (!pre) mbar_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                ::"r"(sa), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(smem_u32(&full_bar[stage])), "r"(kb * GEMM_BK),
                "r"(p.a_row0 + tile_m * GEMM_BM)
                : "memory");
            if (!pre)
              asm volatile(
                  "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                  ::"r"(sa + A_BYTES), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(&full_bar[stage])),
                  "r"(kb * GEMM_BK), "r"(tile_n * BN)
                  : "memory");
          } else {
            const uint32_t lead_full = g2_mapa(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_2cta(sa, &tmA, lead_full, kb * GEMM_BK, p.a_row0 + tile_m * GEMM_BM);
            tma_load_2d_2cta(sa + A_BYTES, &tmB, lead_full, kb * GEMM_BK, tile_n * BN + crank * (BN / CM));
            if (leader) mbar_expect_tx(&full_bar[stage], CM * (A_BYTES + B_BYTES));
          }
          if (tr != nullptr && tile == unit0 && kb == 0) tr[3] = clock64();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      if (tr != nullptr) tr[4] = clock64();
    }
  } else if (warp == 1) {
    // MMA issue: the WHOLE warp runs the loop — barrier waits, stage / descriptor arithmetic are warp-uniform, so the
    // compiler keeps them in uniform registers — and one elected lane issues the tcgen05 instructions.  With the loop under
    // `if (lane == 0)` every tcgen05.mma / commit sat in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall (operands in vector
    // registers of a divergent region): ~65 cycles per MMA, ~670 per 64-wide k-block against 512 of tensor-pipe time for a
    // 128 x 256 tile (tools/mainloop_probe.py, tools/micro/umma_issue.cu).
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM * CM, BN);
      int stage = 0, tcount = 0;
      uint32_t phase = 0, acc = 0, acc_phase = 0;
      for (int tile = unit0; tile < n_tiles; tile += unit_stride) {
        if (CM == 1 && p.tile_needed != nullptr && p.tile_needed[tile / tiles_n] == 0) continue;
        if (tcount >= 2) mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);    // (both accumulators start out free)
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          const bool trk = tr != nullptr && tcount == 0 && kb >= 8 && kb < 12;     // issue-loop timeline, k-blocks 8..11 of the first tile
          if (trk && lane == 0) tr[40 + 5 * (kb - 8)] = clock64();
          mbar_wait(&full_bar[stage], phase);
          if (tr != nullptr && lane == 0 && kb == 0 && tcount < 4) tr[8 + 4 * tcount] = clock64();
          if (trk && lane == 0) tr[41 + 5 * (kb - 8)] = clock64();
          tc_fence_after();
          if (trk && lane == 0) tr[42 + 5 * (kb - 8)] = clock64();
          const uint32_t sa = smem_base + stage * (A_BYTES + B_BYTES);
          const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              if ((dbg & 32) && k > 0) break;   // timing experiment (trace build): one MMA per k-block
              if constexpr (CM == 1) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_bf16_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            if (trk) tr[43 + 5 * (kb - 8)] = clock64();
            if constexpr (CM == 1) umma_commit(&empty_bar[stage]);
            else umma_commit_2cta(&empty_bar[stage]);
            if (trk) tr[44 + 5 * (kb - 8)] = clock64();
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) {
          if constexpr (CM == 1) umma_commit(&tmem_full[acc]);
          else umma_commit_2cta(&tmem_full[acc]);
        }
        __syncwarp();
        if (tr != nullptr && lane == 0 && tcount < 4) tr[9 + 4 * tcount] = clock64();
        ++tcount;
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // 8 epilogue warps: warp w reads TMEM lane quadrant (w & 3) and every second column chunk, starting at `half`.
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t stage_buf = epi_base + (warp - 2) * 4096;   // two 32-row x 64-byte staging tiles per warp
    uint32_t tsel = 0;
    uint32_t acc = 0, acc_phase = 0;
    int tcount = 0;
    for (int tile = unit0; tile < n_tiles; tile += unit_stride) {
      const int tile_m = CM * (tile / tiles_n) + crank, tile_n = tile % tiles_n;
      if (CM == 1 && p.tile_needed != nullptr && p.tile_needed[tile_m] == 0) continue;
      const int m0 = tile_m * GEMM_BM + q * 32;
      const int m = m0 + lane;
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      if constexpr (g2_staged<EPI>()) {
        // a staged 128-byte row holds 32 fp32 (one 32-column sub-chunk) or 64 bf16 (two sub-chunks)
        constexpr int SUB = g2_out_bf16<EPI>() ? 2 : 1;
        constexpr int NCH = BN / (32 * SUB);
        const float* gate = nullptr;
        if constexpr (EPI == EPI_GATE_RES) {
          const int mm = m < p.M ? m : p.M - 1;
          gate = p.mod + static_cast<size_t>(seq_of_row(mm, p.rows_per_utt, p.single)) * p.n_mod + p.gate_off;
        }
        // bias / gate of the NEXT sub-chunk are fetched while the current one is processed (and, for the first,
        // while the accumulator is still being produced): no dependent global latency inside the loop
        // The gate rows are per sequence (L2 latency, measured ~1.5 k cycles under load): they are fetched a whole
        // chunk ahead into a second register set (gqn) at the top of the previous chunk.
        float4 bq[8], gq[EPI == EPI_GATE_RES ? 8 : 1], gqn[EPI == EPI_GATE_RES ? 8 : 1];
        auto prefetch = [&](int col) {
          const int n0 = tile_n * BN + col;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            bq[j] = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        auto prefetch_gate = [&](int col, float4 (&dst)[EPI == EPI_GATE_RES ? 8 : 1]) {
          if constexpr (EPI == EPI_GATE_RES) {
            const int n0 = tile_n * BN + col;
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = __ldg(reinterpret_cast<const float4*>(gate + n0) + j);
          }
        };
        if (half < NCH) { prefetch(half * SUB * 32); prefetch_gate(half * SUB * 32, gq); }
        mbar_wait(&tmem_full[acc], acc_phase);
        if (tr != nullptr && warp == 2 && lane == 0 && tcount < 4) tr[10 + 4 * tcount] = clock64();
        tc_fence_after();
#pragma unroll 1
        for (int c = half; c < NCH; c += 2) {
#pragma unroll
          for (int sub = 0; sub < SUB; ++sub) {
            const int col = (c * SUB + sub) * 32;
            if (EPI == EPI_GATE_RES && c + 2 < NCH) prefetch_gate((c + 2) * SUB * 32, gqn);
            const bool trc = tr != nullptr && warp == 2 && lane == 0 && tcount == 0 && c < 4 && sub == 0;
            long long* te = tr + 32 + (c >> 1) * 8;
            if (trc) te[0] = clock64();
            float v[32];
            {
              uint32_t r[32];
              tmem_ld32(t_addr + col, r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            }
            if (trc) te[1] = clock64();
            if (c + 2 >= NCH && sub == SUB - 1) {  // this warp's last read of the accumulator: hand it back
              tc_fence_before();
              __syncwarp();
              if (lane == 0) { if (CM == 1 || leader) mbar_arrive(&tmem_empty[acc]); else g2_remote_arrive(g2_mapa(smem_u32(&tmem_empty[acc]), 0)); }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[4 * j] += bq[j].x; v[4 * j + 1] += bq[j].y; v[4 * j + 2] += bq[j].z; v[4 * j + 3] += bq[j].w;
            }
            if constexpr (EPI == EPI_F32_POS) {   // + pos[(row / 2) % n_style]  (input projection)
              const int mm = m < p.M ? m : p.M - 1;
              const float4* pr = reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(tok_of_row(mm, p.n_style, p.single)) * p.N + tile_n * BN + col);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 pq = __ldg(pr + j);
                v[4 * j] += pq.x; v[4 * j + 1] += pq.y; v[4 * j + 2] += pq.z; v[4 * j + 3] += pq.w;
              }
            }
            if constexpr (EPI == EPI_GATE_RES) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                v[4 * j] *= gq[j].x; v[4 * j + 1] *= gq[j].y; v[4 * j + 2] *= gq[j].z; v[4 * j + 3] *= gq[j].w;
              }
            }
            if (sub + 1 < SUB) prefetch(col + 32);
            else if (c + 2 < NCH) prefetch((c + 2) * SUB * 32);
            // Two 2 KB staging tiles per warp (32 rows x 64 bytes, 64B-swizzled: 32 bf16 or 16 fp32 columns), used
            // alternately: before a tile is rewritten only the store issued two pieces ago has to have been read out,
            // so the TMA store of one piece overlaps the arithmetic of the next (a single tile serialised every
            // chunk on the ~0.5 us read-out of its predecessor).
            constexpr int PIECES = g2_out_bf16<EPI>() ? 1 : 2;
#pragma unroll
            for (int pc = 0; pc < PIECES; ++pc) {
              if (trc && pc == 0) te[2] = clock64();
              if (elect_one()) bulk_wait_read<1>();
              __syncwarp();
              if (trc && pc == 0) te[3] = clock64();
              const uint32_t tile = stage_buf + tsel * 2048;
              const uint32_t sb = tile + lane * 64, sw = (lane >> 1) & 3;
              if (dbg & 4) {
                if (v[0] == 1234.5f) st_shared_v4(sb, __float_as_uint(v[1]), 0u, 0u, 0u);
              } else if constexpr (EPI == EPI_GELU_BF16) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  st_shared_v4(sb + ((j ^ sw) << 4), gelu_tanh_f16x2_to_bf16x2(v[8 * j], v[8 * j + 1]),
                               gelu_tanh_f16x2_to_bf16x2(v[8 * j + 2], v[8 * j + 3]), gelu_tanh_f16x2_to_bf16x2(v[8 * j + 4], v[8 * j + 5]),
                               gelu_tanh_f16x2_to_bf16x2(v[8 * j + 6], v[8 * j + 7]));
              } else if constexpr (g2_out_bf16<EPI>()) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  st_shared_v4(sb + ((j ^ sw) << 4), pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                               pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  st_shared_v4(sb + ((j ^ sw) << 4), __float_as_uint(v[pc * 16 + 4 * j]), __float_as_uint(v[pc * 16 + 4 * j + 1]),
                               __float_as_uint(v[pc * 16 + 4 * j + 2]), __float_as_uint(v[pc * 16 + 4 * j + 3]));
              }
              if (trc && pc == 0) te[4] = clock64();
              fence_proxy_async();
              __syncwarp();
              if (trc && pc == 0) te[5] = clock64();
              if (elect_one()) {     // (elect.sync region: uniform-register operands, no waterfall around the TMA instruction)
                const int n0 = tile_n * BN + col + pc * 16;
                if (m0 < p.M && !(dbg & 1)) {
                  if constexpr (EPI == EPI_GATE_RES) tma_reduce_add_2d(&tmC, tile, n0, m0);
                  else tma_store_2d(&tmC, tile, n0, m0);
                }
                bulk_commit();
              }
              if (trc && pc == 0) te[6] = clock64();
              tsel ^= 1u;
            }
          }
          if constexpr (EPI == EPI_GATE_RES) {
#pragma unroll
            for (int j = 0; j < 8; ++j) gq[j] = gqn[j];
          }
        }
        if (half >= NCH) {  // (never with NCH >= 2; keeps the barrier count right for any tile shape)
          __syncwarp();
          if (lane == 0) { if (CM == 1 || leader) mbar_arrive(&tmem_empty[acc]); else g2_remote_arrive(g2_mapa(smem_u32(&tmem_empty[acc]), 0)); }
        }
      } else {
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        constexpr int NCH = BN / 32;
#pragma unroll 1
        for (int c = half; c < NCH; c += 2) {
          uint32_t r[32];
          tmem_ld32(t_addr + c * 32, r);
          tmem_ld_wait();
          if (c + 2 >= NCH) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CM == 1 || leader) mbar_arrive(&tmem_empty[acc]); else g2_remote_arrive(g2_mapa(smem_u32(&tmem_empty[acc]), 0)); }
          }
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if constexpr (EPI == EPI_SAMPLER) {
            const int n0 = tile_n * BN + c * 32;
            if (p.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + j);
                v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
              }
            }
            sampler_epilogue32(p, m0, n0, v, stage_buf, lane);
          } else {
            epilogue_row32<EPI>(p, m, tile_n * BN + c * 32, v);
          }
        }
      }
      if (tr != nullptr && warp == 2 && lane == 0 && tcount < 4) tr[11 + 4 * tcount] = clock64();
      ++tcount;
      acc ^= 1u;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (elect_one()) bulk_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (tr != nullptr && threadIdx.x == 0) tr[5] = clock64();
  if constexpr (CM > 1) g2_cluster_sync();   // the peer may still arrive on this CTA's barriers / read its operand stages
  if (warp == 1) {
    if constexpr (CM == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
    else tmem_dealloc2<TMEM_COLS>(tmem_base);
  }
}

}  // namespace stz
