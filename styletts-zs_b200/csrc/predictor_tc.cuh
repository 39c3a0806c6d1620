// lstm_tc_kernel — BiLSTM recurrence (h = 256) with the recurrent matrix-vector product on tcgen05.
//
// lstm_cluster_kernel keeps W_hh in registers and is bound by the fp32 FMA pipe: 128 gate rows x 256 x NB sequences per
// CTA and step = 2.6 k cycles at NB = 10, 3.3 us per time step.  (Legacy mma.sync is no way out: measured
// 565 bf16 FMA/clk/SM on B200, tools/micro/hmma_rate.cu.)  Here the product runs on the 5th-generation tensor core at
// fp32-grade precision with the same split-bf16 scheme as the predictor's input projections:
//     W = W_hi + W_lo,  h = h_hi + h_lo  (bf16 pairs),   W h ~= W_hi h_hi + W_hi h_lo + W_lo h_hi     (error ~2^-16)
// Cluster of 8 CTAs per 16 sequences x direction (as before CTA r owns hidden units [32r, 32r + 32) = 128 gate rows):
//   A operand: the CTA's W_hh slice as two K-major 128B-swizzled bf16 tiles sets (hi, lo; 128 KB), built once per launch;
//   B operand: h_t of the 16 sequences as ONE 32-row tile set (rows 0..15 = h_hi, rows 16..31 = h_lo), double-buffered
//              by step parity: the pass with A = W_hi runs with N = 32 and yields W_hi h_hi and W_hi h_lo in adjacent
//              accumulator columns, the pass with A = W_lo runs with N = 16 over rows 0..15 (tcgen05.mma issue costs
//              ~53 cycles per instruction at these tiny N, so 32 instead of 48 instructions per step matters).  Every CTA
//              writes its 32 new units straight into the swizzled operand tiles of all 8 CTAs with st.async (DSMEM)
//              and signals the destination's mbarrier with the delivered bytes, exactly as the FFMA kernel does;
//   per step : two threads issue 16 tcgen05.mma each (M 128, N 32 | 16, K 16) into two TMEM accumulators and commit;
//              warps 0..3 read it back (thread = gate row), stage the pre-activations in shared memory; all 16 warps
//              then do the pointwise update (warp = sequence, lane = unit), in fp32 as before.
// Packed-sequence semantics as lstm_cluster_kernel (reverse starts at len - 1, padded outputs 0, length-sorted perm).
#pragma once
#include "predictor.cuh"

namespace stz {

// NB = sequences per cluster (8, 16 or 24: every CTA sends its 32 new units of NB sequences to all 8 CTAs each step = NB KB
// out of the SM at the ~17 B/clk DSMEM rate — with W_hh in tensor memory that exchange is the largest part of the step:
// 1.51 k cycles per step at NB = 8, 2.26 k at 16, ~3.0 k at 24.  Only 15 clusters of 8 CTAs are co-resident on a B200,
// so the host picks the NB that minimises waves x step time, see launch_lstm_tc).
constexpr int LT_W_TILE = 128 * 128;                     // one 128-row x 64-k bf16 tile
constexpr int LT_W_BYTES = 2 * 4 * LT_W_TILE;            // hi, lo x 4 k-blocks = 128 KB
constexpr int lt_threads(int nb) { return nb * 32 < 256 ? 256 : nb * 32; }      // warp = sequence; warps 0..5 also drain / issue
constexpr int lt_h_tile(int nb) { return 2 * nb * 128; }                         // one 2 NB-row (hi rows, lo rows) x 64-k bf16 tile
constexpr int lt_h_bytes(int nb) { return 2 * 4 * lt_h_tile(nb); }               // 2 buffers x 4 k-blocks (32 KB at NB = 16)
constexpr int lt_pre_bytes(int nb) { return nb * LC_COLS * 4; }                  // pre-activations [seq][gate row] fp32
constexpr int lt_smem_bytes(bool w_tmem, int nb) { return (w_tmem ? 0 : LT_W_BYTES) + lt_h_bytes(nb) + lt_pre_bytes(nb) + 1024; }
// W_TMEM form: the W_hh slice lives in TENSOR MEMORY as the A operand (tcgen05.mma with A from TMEM): lane = gate row, one
// 32-bit column = two consecutive k (bf16 pair) -> 128 columns for W_hi + 128 for W_lo behind the 64 accumulator columns.
// Every recurrent MMA then reads its A slab from TMEM instead of streaming 4 KB from shared memory (~84 cycles per MMA
// whatever N: the 32 MMAs of a step were 1.34 k of its 3.28 k cycles, profiles/r02_lstm_trace.txt).
constexpr int LT_TMEM_W_HI = 128, LT_TMEM_W_LO = 128 + 128;   // behind the accumulators (3 NB <= 80 columns)

// D[tmem] (+)= A[tmem] * B[smem]^T: A operand from tensor memory (K-major: lane = row, 32-bit column = a k pair)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void st_async_b32(uint32_t remote_addr, uint32_t v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(remote_addr), "r"(v), "r"(remote_bar) : "memory");
}
// Gate non-linearities on the step's critical path: ex2.approx-based (relative error ~1e-7, far inside the split-bf16
// product's 2^-16) instead of the libm expf / tanhf sequences.
__device__ __forceinline__ float lt_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float lt_tanh(float x) {
  const float e = __expf(-2.0f * fabsf(x));               // in (0, 1]: no overflow
  return copysignf(__fdividef(1.0f - e, 1.0f + e), x);
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// Debug timeline (make TRACE=1; tools/lstm_trace.py): cluster 0 / rank 0 records clock64() per step:
// [s][0] step start (warp 4), [1] h landed, [2] MMAs issued, [3] accumulator ready (warp 0), [4] after __syncthreads (warp 0),
// [5] after st.async (warp 0)
__device__ long long* g_lstm_trace = nullptr;

template <bool W_TMEM, int NB>
__global__ void __cluster_dims__(LC_CS, 1, 1) __launch_bounds__(lt_threads(NB), 1)
lstm_tc_kernel(const float* __restrict__ G, const float* __restrict__ Whh, const int* __restrict__ lens,
               const int* __restrict__ perm, float* __restrict__ out, int B, int T) {
  constexpr int LT_NB = NB, LT_THREADS = lt_threads(NB), LT_H_TILE = lt_h_tile(NB), LT_H_BYTES = lt_h_bytes(NB);
  static_assert(NB == 8 || NB == 16 || NB == 24, "operand rows [h_hi ; h_lo] = 2 NB must be a multiple of the 8-row swizzle atom and an MMA N");
  constexpr int N2 = (NB + 15) / 16 * 16;     // N of the W_lo pass (M = 128 needs N % 16 == 0; rows beyond NB are h_lo rows, their columns are never read)
  extern __shared__ uint8_t lt_smem_raw[];
  __shared__ __align__(8) uint64_t hbar[2], mma_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int len_s[LT_NB], seq_s[LT_NB];
  const uint32_t smem_base = (smem_u32(lt_smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t W_SMEM = W_TMEM ? 0u : static_cast<uint32_t>(LT_W_BYTES);
  const uint32_t w_base = smem_base;                       // [hi | lo][k-block][128 rows x 128 B]   (shared-memory form only)
  const uint32_t h_base = smem_base + W_SMEM;              // [buffer][k-block][32 rows x 128 B]: rows 0..15 h_hi, 16..31 h_lo
  float* pre_s = reinterpret_cast<float*>(lt_smem_raw + (smem_base - smem_u32(lt_smem_raw)) + W_SMEM + LT_H_BYTES);   // [seq][gate row]
  constexpr int TMEM_COLS = W_TMEM ? 512 : 64;
  const int rank = static_cast<int>(cluster_ctarank());
  const int group = blockIdx.x / LC_CS, dir = blockIdx.y;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp index provably uniform
#ifdef STZ_TRACE
  long long* tr = (g_lstm_trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) ? g_lstm_trace : nullptr;
#else
  constexpr long long* tr = nullptr;
#endif

  // ---- W_hh slice -> split-bf16 operand tiles (constant: under the previous kernel's tail) ---------------------------
  // gate row r = gate * 32 + unit  <->  W_hh row dir*4h + gate*h + rank*32 + unit; thread -> 8 consecutive k of a row
  for (int i = tid; i < (W_TMEM ? 0 : LC_COLS * (LC_H / 8)); i += LT_THREADS) {
    const int r = i >> 5, c8 = i & 31;                     // row, 8-wide k chunk (32 per row)
    const int g = r >> 5, u = r & 31;
    const float4* src = reinterpret_cast<const float4*>(
        Whh + (static_cast<size_t>(dir) * 4 * LC_H + g * LC_H + rank * LC_UPC + u) * LC_H + c8 * 8);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    uint2 h0, h1;
    h0.x = pack_bf16(a.x, a.y); h0.y = pack_bf16(a.z, a.w);
    h1.x = pack_bf16(b.x, b.y); h1.y = pack_bf16(b.z, b.w);
    const uint2 l0 = split_lo4(a, h0), l1 = split_lo4(b, h1);
    const uint32_t off = (c8 >> 3) * LT_W_TILE + r * 128 + (((c8 & 7) ^ (r & 7)) << 4);
    st_shared_v4(w_base + off, h0.x, h0.y, h1.x, h1.y);
    st_shared_v4(w_base + 4 * LT_W_TILE + off, l0.x, l0.y, l1.x, l1.y);
  }
  for (int i = tid; i < LT_H_BYTES / 16; i += LT_THREADS) st_shared_v4(h_base + i * 16, 0u, 0u, 0u, 0u);   // h_0 = 0
  if (tid == 0) {
    mbar_init(&hbar[0], 1);
    mbar_init(&hbar[1], 1);
    mbar_init(&mma_bar, 2);      // one commit per issuing thread
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(&tmem_slot);
  if constexpr (W_TMEM) {
    // thread = gate row (warps 0..3 = TMEM lane quadrants): the row's 256 weights -> 128 hi words + 128 lo words in TMEM
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp < 4) {
      const int r = tid, g = r >> 5, u = r & 31;
      const float4* src = reinterpret_cast<const float4*>(Whh + (static_cast<size_t>(dir) * 4 * LC_H + g * LC_H + rank * LC_UPC + u) * LC_H);
      const uint32_t lane_base = tmem_slot + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {            // 64 k = 32 packed columns per sweep
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 a = __ldg(src + c * 16 + j);
          uint2 h2;
          h2.x = pack_bf16(a.x, a.y); h2.y = pack_bf16(a.z, a.w);
          const uint2 l2 = split_lo4(a, h2);
          hi[2 * j] = h2.x; hi[2 * j + 1] = h2.y; lo[2 * j] = l2.x; lo[2 * j + 1] = l2.y;
        }
        tmem_st32(lane_base + LT_TMEM_W_HI + c * 32, hi);
        tmem_st32(lane_base + LT_TMEM_W_LO + c * 32, lo);
      }
      tmem_st_wait();
    }
  }
  pdl_sync();
  if (tid < LT_NB) {
    const int idx = group * LT_NB + tid;
    const int seq = idx < B ? perm[idx] : -1;
    seq_s[tid] = seq;
    len_s[tid] = seq >= 0 ? lens[seq] : 0;
  }
  fence_proxy_async();      // operand tiles were written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  int maxlen = 0;
#pragma unroll
  for (int n = 0; n < LT_NB; ++n) maxlen = max(maxlen, len_s[n]);
  for (int n = 0; n < LT_NB; ++n) {  // zero this CTA's unit slice of the padded tail
    if (seq_s[n] < 0) continue;
    for (int t = len_s[n] + warp; t < T; t += LT_THREADS / 32)
      out[(static_cast<size_t>(seq_s[n]) * T + t) * 2 * LC_H + dir * LC_H + rank * LC_UPC + lane] = 0.f;
  }
  cluster_sync_all();   // peers are resident, their barriers initialised and their h tiles zeroed before any DSMEM traffic

  // pointwise update: warp n owns sequence n, lane = unit
  const int my_len = len_s[warp], my_seq = seq_s[warp];
  float c_state = 0.f;
  uint32_t word = 0;      // this lane's packed operand word: even lane (h_hi[u], h_hi[u+1]), odd lane (h_lo[u-1], h_lo[u])
  float gpre[4] = {0.f, 0.f, 0.f, 0.f};
  auto load_g = [&](int s) {
    if (s < my_len) {
      const int t = dir == 0 ? s : my_len - 1 - s;
      const float* gp = G + (static_cast<size_t>(my_seq) * T + t) * 8 * LC_H + dir * 4 * LC_H + rank * LC_UPC + lane;
#pragma unroll
      for (int g = 0; g < 4; ++g) gpre[g] = __ldg(gp + g * LC_H);
    }
  };
  load_g(0);
  // destination of this lane's word inside an h tile set: units k = rank*32 + (lane & ~1), row = sequence `warp`
  const int k0 = rank * LC_UPC + (lane & ~1);
  const uint32_t dst_off = (k0 >> 6) * LT_H_TILE + ((lane & 1) * LT_NB + warp) * 128 +
                           ((((k0 & 63) >> 3) ^ (warp & 7)) << 4) + (k0 & 7) * 2;
  constexpr uint32_t kStepBytes = LT_NB * LC_H * 4;      // hi + lo of 16 x 256 values
  constexpr uint32_t idesc32 = umma_idesc_bf16(128, 2 * LT_NB), idesc16 = umma_idesc_bf16(128, N2);
  constexpr int H_BUF = 4 * LT_H_TILE;                   // bytes per buffer
  // remote addresses of this lane's operand word and of the step barriers in all 8 CTAs (hoisted out of the step loop)
  uint32_t rdst[LC_CS], rbar[LC_CS];
#pragma unroll
  for (int r = 0; r < LC_CS; ++r) {
    rdst[r] = mapa_u32(h_base + dst_off, r);
    rbar[r] = mapa_u32(smem_u32(&hbar[0]), r);
  }
  const uint32_t bar_step = smem_u32(&hbar[1]) - smem_u32(&hbar[0]);

  for (int s = 0; s < maxlen; ++s) {
    const int cur = s & 1, nxt = cur ^ 1;
    if (tid == 0) mbar_expect_tx(&hbar[nxt], kStepBytes);   // arm the buffer this step's h will land in
    if (warp == 4 || warp == 5) {
      // two issuing warps: warp 4: W_hi x [h_hi ; h_lo] (N = 32) -> columns 0..31;  warp 5: W_lo x h_hi (N = 16) -> columns 32..47.
      // The whole warp runs the (warp-uniform) wait and descriptor arithmetic and ONE elected lane issues: under
      // `lane == 0` every tcgen05.mma sat in an ELECT / R2UR.BROADCAST waterfall, ~65 cycles each instead of ~28
      // (tools/micro/umma_issue.cu) — 16 of them head every time step's dependency chain.
      const int seg = warp - 4;
      if (tr != nullptr && seg == 0 && s < 64) tr[s * 8 + 0] = clock64();
      if (s > 0) {
        const uint32_t parity = ((s - 1) >> 1) & 1;
        uint32_t spins = 0;
        while (!mbar_try_wait_cluster(&hbar[cur], parity)) {
          if (++spins > (1u << 24)) __trap();
        }
      }
      if (tr != nullptr && seg == 0 && s < 64) tr[s * 8 + 1] = clock64();
      fence_proxy_async();
      tc_fence_after();
      const uint32_t hb = h_base + cur * H_BUF, wa = w_base + seg * 4 * LT_W_TILE, dcol = tmem_d + seg * 2 * LT_NB;
      const uint32_t wt = tmem_d + (seg == 0 ? LT_TMEM_W_HI : LT_TMEM_W_LO);     // W_TMEM: 8 columns per K = 16 slab
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          const uint64_t db = umma_desc_sw128(hb + kb * LT_H_TILE);
          if constexpr (W_TMEM) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ts(dcol, wt + (kb * 4 + k) * 8, db + 2 * k, seg == 0 ? idesc32 : idesc16, (kb | k) != 0 ? 1u : 0u);
          } else {
            const uint64_t da = umma_desc_sw128(wa + kb * LT_W_TILE);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(dcol, da + 2 * k, db + 2 * k, seg == 0 ? idesc32 : idesc16, (kb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&mma_bar);
      }
      __syncwarp();
      if (tr != nullptr && seg == 0 && s < 64) tr[s * 8 + 2] = clock64();
    }
    if (warp < 4) {     // thread = gate row: accumulator -> pre_s[seq][gate row]
      mbar_wait(&mma_bar, s & 1);
      if (tr != nullptr && warp == 0 && s < 64) tr[s * 8 + 3] = clock64();
      tc_fence_after();
      // accumulator columns: [0, NB) W_hi h_hi, [NB, 2 NB) W_hi h_lo, [2 NB, 3 NB) W_lo h_hi
      const uint32_t lane_base = tmem_d + (static_cast<uint32_t>(warp * 32) << 16);
      uint32_t a[32], b[NB > 8 ? (NB > 16 ? 32 : 16) : 1], c[NB > 16 ? 16 : 1];
      tmem_ld32(lane_base, a);
      if constexpr (NB == 16) tmem_ld16(lane_base + 32, b);
      if constexpr (NB == 24) { tmem_ld32(lane_base + 32, b); tmem_ld16(lane_base + 64, c); }
      tmem_ld_wait();
      auto col = [&](int i) -> float { return __uint_as_float(i < 32 ? a[i & 31] : (i < 64 ? b[(i & 31) % (NB > 16 ? 32 : 16)] : c[(i & 15) % (NB > 16 ? 16 : 1)])); };
#pragma unroll
      for (int n = 0; n < LT_NB; ++n) pre_s[n * LC_COLS + tid] = (col(n) + col(LT_NB + n)) + col(2 * LT_NB + n);
      tc_fence_before();
    }
    __syncthreads();
    if (tr != nullptr && warp == 0 && s < 64) tr[s * 8 + 4] = clock64();
    {
      float hn = 0.f;
      const bool active = s < my_len;
      if (active) {
        const float* ps = pre_s + warp * LC_COLS + lane;
        const float ig = lt_sigmoid(gpre[0] + ps[0]), fg = lt_sigmoid(gpre[1] + ps[32]);
        const float gt = lt_tanh(gpre[2] + ps[64]), og = lt_sigmoid(gpre[3] + ps[96]);
        c_state = fg * c_state + ig * gt;
        hn = og * lt_tanh(c_state);
      }
      // split and pair up: even lane carries the hi pair, odd lane the lo pair of units (lane & ~1, lane | 1)
      const __nv_bfloat16 hi = __float2bfloat16(hn);
      const __nv_bfloat16 lo = __float2bfloat16(hn - __bfloat162float(hi));
      const uint32_t mine = (static_cast<uint32_t>(__bfloat16_as_ushort(lo)) << 16) | __bfloat16_as_ushort(hi);
      const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
      if (active) {
        word = (lane & 1) ? ((mine & 0xffff0000u) | (other >> 16))            // (lo[u-1], lo[u])
                          : ((other << 16) | (mine & 0xffffu));               // (hi[u], hi[u+1])
      }
#pragma unroll
      for (int r = 0; r < LC_CS; ++r) st_async_b32(rdst[r] + nxt * H_BUF, word, rbar[r] + nxt * bar_step);
      if (tr != nullptr && warp == 0 && s < 64) tr[s * 8 + 5] = clock64();
      if (active) {
        const int t = dir == 0 ? s : my_len - 1 - s;
        out[(static_cast<size_t>(my_seq) * T + t) * 2 * LC_H + dir * LC_H + rank * LC_UPC + lane] = hn;
      }
    }
    load_g(s + 1);
  }
  // every CTA must have received (and not be waiting for) the last step's stores before anyone exits
  if (maxlen > 0 && warp == 4 && lane == 0) {   // (one waiter is enough: the barrier phase is CTA-wide state)
    const uint32_t parity = ((maxlen - 1) >> 1) & 1;
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(&hbar[maxlen & 1], parity)) {
      if (++spins > (1u << 24)) __trap();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem_d);
}

}  // namespace stz
