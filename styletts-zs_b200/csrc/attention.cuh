// Fused attention over the fixed-length style-code sequence.
//
// One CTA per (utterance b, head): the 2*K query rows of utterance b (both CFG branches, R layout,
// contiguous) against a "virtual" key sequence assembled from up to three segments:
//   self-attention : the utterance's own 2*K rows, a key is visible to queries of the same branch;
//   cross-attention: [text keys (shared by both branches, padding-masked) ; prompt keys (cond branch
//                    only, masked) ; the null-prompt key (uncond branch only)].
// Flash-style streaming over 64-key blocks with an fp32 online softmax; QK^T and PV run on the
// legacy warp-level tensor path (mma.sync m16n8k16 bf16 -> fp32).  This is ~4 % of the
// denoiser's flops; the tcgen05 budget goes to the GEMMs (gemm.cuh).
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace stz {

enum KeyRule : int { KEY_ALL = 0, KEY_COND = 1, KEY_UNCOND = 2, KEY_SAME_BRANCH = 3 };

struct AttnSeg {
  const __nv_bfloat16* k;  // first key row of utterance 0 (already offset to this layer / K or V columns: see k_col)
  const __nv_bfloat16* v;
  int ld;                  // row stride in elements
  int n;                   // keys per utterance in this segment
  int rows_per_utt;        // row advance per utterance (0 = shared by all utterances)
  const uint8_t* mask;     // [B, n] 1 = valid, or nullptr
  int rule;
};

struct AttnParams {
  const __nv_bfloat16* q;  // R layout, head h at columns h*64
  int ldq;
  __nv_bfloat16* out;      // R layout [R, ldo]
  int ldo;
  int n_q;                 // 2 * n_style (<= 128)
  int nseg;
  AttnSeg seg[3];
  float scale_log2;        // log2(e) / sqrt(d_head)
  int single;              // single-branch row layout: n_q = n_style query rows per utterance (tcgen05 kernels only)
};

constexpr int ATT_DH = 64;
constexpr int ATT_KB = 64;       // keys per block
constexpr int ATT_LDS = 72;      // padded smem row (bf16 elements): 144 B, conflict-free for the fragment loads

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
// 16-byte async copy global -> shared; src_bytes = 0 zero-fills (masked / out-of-range keys)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int ATT_MAXQ = 128;
constexpr int ATT_SMEM_BYTES = (ATT_MAXQ + 4 * ATT_KB) * ATT_LDS * 2 + 2 * ATT_KB;   // Q + 2 x (K, V) + visibility

// Warps are split by CFG branch: warp w handles 16 style tokens of branch w / nwb (nwb = ceil(K / 16)), so a warp
// only ever multiplies against keys its branch can see:
//   self-attention : keys are de-interleaved at load time, block 0 = conditional rows, block 1 = unconditional
//                    rows; a warp processes exactly one 64-key block (instead of 2 x 64 half-masked keys);
//   cross-attention: 8-key tiles without a visible key for the warp's branch are skipped (warp-uniform predicate).
// All global loads are asynchronous (cp.async): Q and the first two key blocks are in flight before any math,
// later blocks are fetched into the buffer just consumed (double buffering).
__global__ void __launch_bounds__(256, 2) attention_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  __nv_bfloat16 (*Qs)[ATT_LDS] = reinterpret_cast<__nv_bfloat16 (*)[ATT_LDS]>(att_smem);
  __nv_bfloat16 (*KVs)[ATT_LDS] = reinterpret_cast<__nv_bfloat16 (*)[ATT_LDS]>(att_smem + ATT_MAXQ * ATT_LDS * 2);
  // KVs rows: [buf][0: K | 1: V][ATT_KB]
  uint8_t* kvis = att_smem + (ATT_MAXQ + 4 * ATT_KB) * ATT_LDS * 2;   // [2][ATT_KB]; bit0: cond rows, bit1: uncond rows
  pdl_sync();

  const int head = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n_tok = p.n_q >> 1;                       // style tokens per branch
  const int nwb = (blockDim.x >> 5) >> 1;             // warps per branch
  const int br = warp / nwb, qt = warp - br * nwb;
  const int tok0 = qt * 16 + g, tok1 = tok0 + 8;      // this thread's two style tokens (rows of the MMA tile)
  const size_t qbase = static_cast<size_t>(b) * p.n_q;
  const bool self_attn = p.nseg == 1 && p.seg[0].rule == KEY_SAME_BRANCH;

  int n_total = 0;
  for (int s = 0; s < p.nseg; ++s) n_total += p.seg[s].n;
  const int n_blk = self_attn ? 2 : (n_total + ATT_KB - 1) / ATT_KB;

  // one thread = one key row, two 16-byte chunks of K and of V (blockDim.x == 256: 64 rows x 4 threads)
  auto load_block = [&](int blk, int buf) {
    for (int i = threadIdx.x; i < ATT_KB * 4; i += blockDim.x) {
      const int kr = i >> 2, c0 = (i & 3) * 2;
      const __nv_bfloat16 *ksrc = p.seg[0].k, *vsrc = p.seg[0].v;
      uint32_t bytes = 0;
      uint8_t vis = 0;
      int vk = self_attn ? 2 * kr + blk : blk * ATT_KB + kr;   // self: block = branch, slot = token
      if (vk < n_total && (!self_attn || kr < n_tok)) {
        int s = 0;
        while (vk >= p.seg[s].n) { vk -= p.seg[s].n; ++s; }
        const AttnSeg& sg = p.seg[s];
        const bool ok = sg.mask == nullptr || sg.mask[static_cast<size_t>(b) * sg.n + vk] != 0;
        if (ok) {
          const size_t r = static_cast<size_t>(b) * sg.rows_per_utt + vk;
          ksrc = sg.k + r * sg.ld + head * ATT_DH;
          vsrc = sg.v + r * sg.ld + head * ATT_DH;
          bytes = 16;
          vis = sg.rule == KEY_ALL ? 3 : sg.rule == KEY_COND ? 1 : sg.rule == KEY_UNCOND ? 2 : ((vk & 1) ? 2 : 1);
        }
      }
#pragma unroll
      for (int c = c0; c < c0 + 2; ++c) {
        cp_async16(smem_u32(&KVs[(buf * 2 + 0) * ATT_KB + kr][c * 8]), ksrc + c * 8, bytes);
        cp_async16(smem_u32(&KVs[(buf * 2 + 1) * ATT_KB + kr][c * 8]), vsrc + c * 8, bytes);
      }
      if (c0 == 0) kvis[buf * ATT_KB + kr] = vis;
    }
  };

  // Q tile, de-interleaved: smem row (br * nwb + qt) * 16 + i  <-  global row 2 * tok + br   (tok >= n_tok: zeros)
  for (int i = threadIdx.x; i < (blockDim.x >> 5) * 16 * 8; i += blockDim.x) {
    const int r = i >> 3, ch = i & 7;
    const int rb = r / (nwb * 16), tok = r - rb * nwb * 16;
    const bool ok = tok < n_tok;
    cp_async16(smem_u32(&Qs[r][ch * 8]), p.q + (qbase + (ok ? 2 * tok + rb : 0)) * p.ldq + head * ATT_DH + ch * 8, ok ? 16u : 0u);
  }
  load_block(0, 0);
  cp_async_commit();
  if (n_blk > 1) load_block(1, 1);
  cp_async_commit();

  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  uint32_t qa[4][4];
  auto load_q = [&]() {   // Q fragments (A operand, 16 x 64 per warp) via ldmatrix
    const int mi = lane >> 3, r = lane & 7;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      ldmatrix_x4(qa[kk], smem_u32(&Qs[warp * 16 + (mi & 1) * 8 + r][kk * 16 + (mi >> 1) * 8]));
  };

  auto process = [&](int buf) {
    const __nv_bfloat16 (*Ks)[ATT_LDS] = &KVs[(buf * 2 + 0) * ATT_KB];
    const __nv_bfloat16 (*Vs)[ATT_LDS] = &KVs[(buf * 2 + 1) * ATT_KB];
    const uint8_t* kv = kvis + buf * ATT_KB;
    // 8-key tiles with at least one key visible to this warp's branch
    const uint32_t pair = __ballot_sync(0xffffffffu, (((kv[2 * lane] | kv[2 * lane + 1]) >> br) & 1) != 0);
    uint32_t tmask = 0;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) tmask |= ((pair >> (4 * nt)) & 0xFu) ? (1u << nt) : 0u;
    if (tmask == 0) return;

    // S = Q K^T  (16 x 64 per warp), invisible tiles skipped
    float sc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
      if ((tmask >> nt) & 1) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Ks[nt * 8 + g][kk * 16 + t * 2]);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&Ks[nt * 8 + g][kk * 16 + 8 + t * 2]);
          mma_bf16_16816(sc[nt], qa[kk], b0, b1);
        }
      }
    }
    // scale, mask, block row max (both rows of this thread belong to branch br)
    float bm[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const bool vis = ((tmask >> nt) & 1) && ((kv[nt * 8 + t * 2 + j] >> br) & 1);
        sc[nt][j] = vis ? sc[nt][j] * p.scale_log2 : -INFINITY;
        sc[nt][2 + j] = vis ? sc[nt][2 + j] * p.scale_log2 : -INFINITY;
        bm[0] = fmaxf(bm[0], sc[nt][j]);
        bm[1] = fmaxf(bm[1], sc[nt][2 + j]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      bm[r] = fmaxf(bm[r], __shfl_xor_sync(0xffffffffu, bm[r], 1));
      bm[r] = fmaxf(bm[r], __shfl_xor_sync(0xffffffffu, bm[r], 2));
    }
    float alpha[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mnew[r] = fmaxf(mrow[r], bm[r]);
      const float base = mnew[r] == -INFINITY ? 0.f : mnew[r];
      alpha[r] = exp2f(mrow[r] - base);  // mrow = -inf -> 0
      mrow[r] = mnew[r];
      mnew[r] = base;
      lrow[r] *= alpha[r];
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      o[nt][0] *= alpha[0]; o[nt][1] *= alpha[0]; o[nt][2] *= alpha[1]; o[nt][3] *= alpha[1];
      if ((tmask >> nt) & 1) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          sc[nt][j] = exp2f(sc[nt][j] - mnew[0]);
          sc[nt][2 + j] = exp2f(sc[nt][2 + j] - mnew[1]);
          lrow[0] += sc[nt][j];
          lrow[1] += sc[nt][2 + j];
        }
      } else {
        sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
      }
    }
    // O += P V  (P rounded to bf16 as the MMA A operand; row sums stay fp32); 16-key steps without visible keys skipped
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (((tmask >> (2 * kk)) & 3u) == 0) continue;
      uint32_t pa[4];
      pa[0] = pack_bf16(sc[2 * kk][0], sc[2 * kk][1]);
      pa[1] = pack_bf16(sc[2 * kk][2], sc[2 * kk][3]);
      pa[2] = pack_bf16(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
      pa[3] = pack_bf16(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
      for (int ntp = 0; ntp < 4; ++ntp) {  // two 8-wide dim tiles per ldmatrix.x4
        const int mi = lane >> 3, r = lane & 7;
        const uint32_t addr = smem_u32(&Vs[kk * 16 + (mi & 1) * 8 + r][ntp * 16 + (mi >> 1) * 8]);
        uint32_t vb[4];
        ldmatrix_x4_trans(vb, addr);
        mma_bf16_16816(o[2 * ntp], pa, vb[0], vb[1]);
        mma_bf16_16816(o[2 * ntp + 1], pa, vb[2], vb[3]);
      }
    }
  };

  if (self_attn) {
    cp_async_wait<0>();
    __syncthreads();
    load_q();
    process(br);            // block br holds exactly the keys of this warp's branch
  } else {
    for (int blk = 0; blk < n_blk; ++blk) {
      const int buf = blk & 1;
      cp_async_wait<1>();      // this block's group (and Q) has landed; the next block may still be in flight
      __syncthreads();
      if (blk == 0) load_q();
      process(buf);
      __syncthreads();                       // everyone is done with buffer `buf`
      if (blk + 2 < n_blk) load_block(blk + 2, buf);
      cp_async_commit();                     // (possibly empty) keeps the group count in step with the block index
    }
  }
  // finalize
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
  }
  const float inv0 = lrow[0] > 0.f ? 1.f / lrow[0] : 0.f, inv1 = lrow[1] > 0.f ? 1.f / lrow[1] : 0.f;
  if (tok0 < n_tok) {
    __nv_bfloat16* op = p.out + (qbase + 2 * tok0 + br) * p.ldo + head * ATT_DH + t * 2;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<uint32_t*>(op + nt * 8) = pack_bf16(o[nt][0] * inv0, o[nt][1] * inv0);
  }
  if (tok1 < n_tok) {
    __nv_bfloat16* op = p.out + (qbase + 2 * tok1 + br) * p.ldo + head * ATT_DH + t * 2;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<uint32_t*>(op + nt * 8) = pack_bf16(o[nt][2] * inv1, o[nt][3] * inv1);
  }
}

// ------------------------------------------------------------------------------------------------------
// tcgen05 resident-key attention (n_total <= 128, n_q <= 128): shared scheme of attention_tc2_kernel / attention_tcs_kernel.
//
// The mma.sync kernels above are instruction-issue bound (~9 k warp instructions per (utterance, head) unit,
// 64 of them HMMA).  Here both contractions run on the 5th-generation tensor core and the softmax is
// thread-per-row, with no shuffles:
//   S[128 q x 128 keys] = Q K^T      4 x tcgen05.mma (M 128, N 128, K 16), operands K-major, 128B-swizzled
//   P = softmax(S) (masked)          thread r owns TMEM lane r = query row r: two passes over its S row with
//                                    tcgen05.ld (max, then exp2 / sum), P written as bf16 into shared memory in
//                                    the K-major swizzled layout (invisible keys = 0), over the dead Q | K tiles
//   O[128 q x 64] = P V              8 x tcgen05.mma (M 128, N 64, K 16); V stays [key][d_head] as loaded and is
//                                    consumed as an MN-major B operand (no transpose)
//   out = O / rowsum                 tcgen05.ld, one 128-byte row per thread
// Row r of the tile = branch (r >> 6), style token (r & 63): a warp is branch-uniform.  Persistent CTAs of 4 warps,
// 2 per SM; the next unit's Q/K/V are fetched while the current unit is computed.
// ------------------------------------------------------------------------------------------------------
constexpr int ATC_TILE = 128 * 128;                 // one 128-row x 128-byte operand tile
constexpr int ATC_BUF_BYTES = 3 * ATC_TILE;         // Q | K | V   (P later overwrites Q | K)
constexpr int ATC_SMEM_BYTES = 2 * ATC_BUF_BYTES + 1024;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// MN-major B operand: rows = 8-deep K groups (here: keys), 128 B (64 bf16 of the N dimension) per K index,
// 128B swizzle, groups of 8 K indices 1024 B apart (SBO); a single 64-wide N atom, so LBO is unused.
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1024 >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Debug timeline (tools/att_trace.py): when set, thread 0 of every CTA records clock64() at the phase boundaries of
// its first two units into g_att_trace[cta][16].
__device__ long long* g_att_trace = nullptr;

// ------------------------------------------------------------------------------------------------------
// attention_tc2_kernel — the scheme above with (a) TMA operand staging and (b) 8 warps.
//
// (a) The timeline of a cp.async-staged first version (tools/att_trace.py) showed ~1.5 k cycles of LSU time per unit just
//     issuing the 3072 16-byte cp.async of one unit.  Here one thread issues 6-8 TMA box copies per unit:
//       self  : a 3-D view (column, branch, token) of the qkv buffer de-interleaves the CFG branches in the copy:
//               box (64 cols, 1 branch, 64 tokens) -> tile rows branch * 64 + token, for Q, K and V;
//       cross : Q as above; text / prompt / null-prompt K and V as 2-D boxes at tile rows 0, T8 and T8 + P8
//               (segments start on 8-row = 1024-byte swizzle-atom boundaries).
//     Rows a box does not cover keep old (finite: the tiles are zeroed at kernel start) contents and are masked by
//     the per-key visibility byte; padded / masked keys are loaded but invisible (their K/V rows are finite because
//     cast_pool_kernel zeroes masked tokens).
// (b) Warps w and w + 4 share TMEM lane quadrant w & 3 (= the same 32 query rows) and split the 32-key chunks of
//     S by parity and the 64 output columns in halves; row max / row sum partials meet in shared memory.
//     Fully visible chunks skip the per-element visibility select (warp-uniform branch).
// ------------------------------------------------------------------------------------------------------
struct AttnTcParams {
  __nv_bfloat16* out;      // R layout [R, ldo]
  int ldo, n_q, n_heads, n_units;
  int self;                // 1: keys = the utterance's own rows (same branch); 0: [text ; prompt | null]
  int T, P, T8, P8;        // cross: segment lengths and their 8-row padded sizes
  int col_k, col_v;        // first column of this layer's K / V inside the K/V tensor maps (head 0)
  const uint8_t* tmask;    // [B, T] or nullptr
  const uint8_t* pmask;    // [B, P] or nullptr
  int n_style;             // tokens per branch
  float scale_log2;
  int single;              // single-branch row layout (guidance-conditioned student): query row = b * n_style + token, only tile
                           // rows 0..63 (the "conditional" half) are loaded and written; the unit is still (utterance, head)
  int box2;                // 2: self-attention, tmT is a 4-D (column, token, branch, part) view: ONE box lands Q | K | V;
                           // 1: the Q (self: Q / K / V) tensor map is (column, token, branch) with a box of 64 x 64 x branches: ONE
                           // TMA instruction per operand lands both branches de-interleaved (TMA issues serialise at ~200
                           // cycles each per SM: tools/att_trace.py); 0: (column, branch, token) map, one box per branch
};

// utterance index of an (utterance, head) unit: a shift when the head count is a power of two (an integer division on the
// uniform datapath is a ~250-cycle dependent chain, and it sits in front of every unit's first TMA issue)
__device__ __forceinline__ int unit_utt(int unit, int n_heads) {
  return (n_heads & (n_heads - 1)) == 0 ? unit >> (31 - __clz(n_heads)) : unit / n_heads;
}

__device__ __forceinline__ void tma_load_2d_u32(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d_u32(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void tma_load_4d_u32(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__global__ void __launch_bounds__(256, 2) attention_tc2_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                               const __grid_constant__ CUtensorMap tmT,
                                                               const __grid_constant__ CUtensorMap tmP,
                                                               const __grid_constant__ CUtensorMap tmN,
                                                               const AttnTcParams p) {
  extern __shared__ uint8_t atc_smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[2], bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  __shared__ uint8_t visb[2][128];
  __shared__ uint32_t colmask[2][4];        // [branch][32-column chunk] of the current unit
  __shared__ float pmax[2][128], psum[2][128];
  const uint32_t smem_base = (smem_u32(atc_smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp index provably uniform
  const int q4 = warp & 3, half = warp >> 2;
  const int row = q4 * 32 + lane;                      // query row = TMEM lane
  const int br = row >> 6;                             // rows 0..63 conditional, 64..127 unconditional
  const int n_tok = p.n_style;
#ifdef STZ_TRACE
  long long* tr = (g_att_trace != nullptr && tid == 0) ? g_att_trace + blockIdx.x * 32 : nullptr;
#else
  constexpr long long* tr = nullptr;
#endif
  int tri = 0;
#define ATC2_TR() do { if (tr != nullptr && tri < 32) tr[tri++] = clock64(); } while (0)
  ATC2_TR();

  if (tid == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmT);
    if (!p.self) { prefetch_tmap(&tmP); prefetch_tmap(&tmN); }
    mbar_init(&bar_full[0], 1);
    mbar_init(&bar_full[1], 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(&tmem_slot);
  // Only V rows that no TMA box covers must be finite (p = 0 times NaN would poison O); K rows the boxes do not cover only
  // produce S columns that the visibility select discards, and Q / P rows are always fully written.  Self-attention boxes
  // cover every row, so nothing is cleared there; cross-attention clears the two V tiles.
  if (!p.self || p.single) {     // (single-branch self-attention loads only rows 0..63 of V)
    for (uint32_t o = tid * 16; o < 2 * ATC_TILE; o += 256 * 16) {
      const uint32_t buf = o / ATC_TILE, off = o % ATC_TILE;
      st_shared_v4(smem_base + buf * ATC_BUF_BYTES + 2 * ATC_TILE + off, 0u, 0u, 0u, 0u);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = tmem_slot, tmem_o = tmem_slot + 128;
  ATC2_TR();
  const uint32_t nbox = p.single ? 1u : 2u;             // 64-row boxes per operand (one per branch)
  const uint32_t tx_bytes = p.self ? 3u * nbox * 8192u : nbox * 8192u + 2u * 128u * static_cast<uint32_t>(p.T + p.P + 1);
  // parts: 1 = Q (written by the previous kernel), 2 = K / V.  The cross-attention K / V of a call are constants of the
  // evaluation loop (computed once in the conditioning prep), so the first unit's K / V boxes are requested BEFORE
  // griddepcontrol.wait and land while the previous kernel drains.
  // Called by lane 0 of EVERY warp: each warp issues one box (a TMA issue costs ~150 cycles; eight of them from one
  // thread held the whole CTA back at the next barrier).  Warp 0 arms the barrier with the unit's total bytes; boxes of
  // other warps may complete before that — the transaction count is signed and the phase cannot complete before warp
  // 0's arrive.
  auto produce_tma = [&](int unit, int buf, int parts) {
    const int b = unit_utt(unit, p.n_heads), head = unit - b * p.n_heads;
    const uint32_t qs = smem_base + buf * ATC_BUF_BYTES, ks = qs + ATC_TILE, vs = ks + ATC_TILE;
    const uint32_t bar = smem_u32(&bar_full[buf]);
    const int t0 = b * n_tok, hc = head * ATT_DH;
    // (every lane computes the warp-uniform addresses above; ONE elected lane issues — inside an elect.sync region the compiler
    // may keep operands in uniform registers, under `lane == 0` every TMA instruction sat in an ELECT / R2UR waterfall)
    if (!elect_one()) return;
    if (parts & 2) {
      if (warp == 0) mbar_expect_tx(&bar_full[buf], tx_bytes);   // the whole unit's bytes, armed with its first part
      if (p.self) {
        if (p.box2 == 2) {         // Q | K | V tiles in one 4-D box (tmT: column, token, branch, part)
          if (warp == 0) tma_load_4d_u32(qs, &tmT, bar, hc, t0, 0, 0);
        } else if (p.box2) {
          if (warp == 0) tma_load_3d_u32(ks, &tmQ, bar, p.col_k + hc, t0, 0);
          else if (warp == 2) tma_load_3d_u32(vs, &tmQ, bar, p.col_v + hc, t0, 0);
        } else {
          if (warp == 0) tma_load_3d_u32(ks, &tmQ, bar, p.col_k + hc, 0, t0);
          else if (warp == 1 && !p.single) tma_load_3d_u32(ks + 8192, &tmQ, bar, p.col_k + hc, 1, t0);
          else if (warp == 2) tma_load_3d_u32(vs, &tmQ, bar, p.col_v + hc, 0, t0);
          else if (warp == 3 && !p.single) tma_load_3d_u32(vs + 8192, &tmQ, bar, p.col_v + hc, 1, t0);
        }
      } else {
        const uint32_t o1 = p.T8 * 128, o2 = (p.T8 + p.P8) * 128;
        if (warp == 0) tma_load_2d_u32(ks, &tmT, bar, p.col_k + hc, b * p.T);
        else if (warp == 1) tma_load_2d_u32(vs, &tmT, bar, p.col_v + hc, b * p.T);
        else if (warp == 2) tma_load_2d_u32(ks + o1, &tmP, bar, p.col_k + hc, b * p.P);
        else if (warp == 3) tma_load_2d_u32(vs + o1, &tmP, bar, p.col_v + hc, b * p.P);
        else if (warp == 4) tma_load_2d_u32(ks + o2, &tmN, bar, p.col_k + hc, 0);
        else if (warp == 5) tma_load_2d_u32(vs + o2, &tmN, bar, p.col_v + hc, 0);
      }
    }
    if (parts & 1) {
      if (p.box2 == 2 && p.self) {
      } else if (p.box2) {
        if (warp == 6) tma_load_3d_u32(qs, &tmQ, bar, hc, t0, 0);
      } else {
        if (warp == 6) tma_load_3d_u32(qs, &tmQ, bar, hc, 0, t0);
        else if (warp == 7 && !p.single) tma_load_3d_u32(qs + 8192, &tmQ, bar, hc, 1, t0);
      }
    }
  };
  const bool kv_early = !p.self && static_cast<int>(blockIdx.x) < p.n_units;
  if (kv_early) produce_tma(blockIdx.x, 0, 2);
  pdl_sync();
  ATC2_TR();

  auto produce = [&](int unit, int buf, int parts) {
    const int b = unit_utt(unit, p.n_heads);
    produce_tma(unit, buf, parts);
    if (tid < 128) {   // visibility of key row `tid`: bit 0 = conditional queries, bit 1 = unconditional queries
      uint8_t vis = 0;
      if (p.self) {
        vis = (tid & 63) < n_tok ? static_cast<uint8_t>(1u << (tid >> 6)) : 0;
      } else if (tid < p.T) {
        vis = (p.tmask == nullptr || p.tmask[static_cast<size_t>(b) * p.T + tid] != 0) ? 3 : 0;
      } else if (tid >= p.T8 && tid < p.T8 + p.P) {
        vis = (p.pmask == nullptr || p.pmask[static_cast<size_t>(b) * p.P + (tid - p.T8)] != 0) ? 1 : 0;
      } else if (tid == p.T8 + p.P8) {
        vis = 2;
      }
      visb[buf][tid] = vis;
    }
  };

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64) | (1u << 16);   // B (= V) is MN-major
  const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
  int buf = 0;
  uint32_t phase = 0, fph0 = 0, fph1 = 0;
  if (static_cast<int>(blockIdx.x) < p.n_units) produce(blockIdx.x, 0, kv_early ? 1 : 3);
  for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
    const int next = unit + gridDim.x;
    if (next < p.n_units) produce(next, buf ^ 1, 3);
    ATC2_TR();
    mbar_wait(&bar_full[buf], buf ? fph1 : fph0);
    if (buf) fph1 ^= 1u; else fph0 ^= 1u;
    __syncthreads();              // visb[buf] visible
    ATC2_TR();
    const uint32_t qs = smem_base + buf * ATC_BUF_BYTES, ks = qs + ATC_TILE, vs = ks + ATC_TILE;
    if (warp == 0) {     // warp-uniform operands, one elected lane issues (no per-instruction waterfall: see gemm2_kernel)
      tc_fence_after();
      const uint64_t da = umma_desc_sw128(qs), db = umma_desc_sw128(ks);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_s, da + 2 * k, db + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&bar_s);
      }
      __syncwarp();
    }
    if (warp < 4) {  // per-branch visibility of each 32-key chunk
      const uint8_t v = visb[buf][warp * 32 + lane];
      const uint32_t m0 = __ballot_sync(0xffffffffu, (v & 1) != 0), m1 = __ballot_sync(0xffffffffu, (v & 2) != 0);
      if (lane == 0) { colmask[0][warp] = m0; colmask[1][warp] = m1; }
    }
    __syncthreads();              // colmask visible
    const uint32_t cm0 = colmask[br][half], cm1 = colmask[br][half + 2];   // this warp's chunks: half, half + 2
    mbar_wait(&bar_s, phase);
    tc_fence_after();
    ATC2_TR();

    // ---- pass 1: row max over the visible keys of this warp's chunks
    float mx = -INFINITY;
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) {
      const uint32_t cm = ci ? cm1 : cm0;
      if (cm == 0) continue;   // warp-uniform
      uint32_t r[32];
      tmem_ld32(tmem_s + lane_addr + (half + 2 * ci) * 32, r);
      tmem_ld_wait();
      if (cm == 0xffffffffu) {   // four independent chains (a single running max is a 16-deep dependent FMNMX chain per chunk)
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 32; j += 8)
#pragma unroll
          for (int i = 0; i < 4; ++i) m4[i] = fmaxf(m4[i], fmaxf(__uint_as_float(r[j + 2 * i]), __uint_as_float(r[j + 2 * i + 1])));
        mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
      } else {   // partially visible chunk: mask, so that the result never depends on a neighbouring utterance's rows
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, ((cm >> j) & 1u) ? __uint_as_float(r[j]) : -INFINITY);
      }
    }
    pmax[half][row] = mx;
    __syncthreads();
    ATC2_TR();
    mx = fmaxf(pmax[0][row], pmax[1][row]);
    const float mb = (mx == -INFINITY ? 0.f : mx) * p.scale_log2;

    // ---- pass 2: P = exp2(S * scale - max), masked, bf16, into the dead Q | K tiles (K-major, swizzled)
    float lsum = 0.f;
    const uint32_t prow = qs + row * 128, sw = row & 7;
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) {
      const int c = half + 2 * ci;
      const uint32_t cm = ci ? cm1 : cm0;
      const uint32_t pb = prow + (c >> 1) * ATC_TILE;
      if (cm == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) st_shared_v4(pb + ((((c & 1) * 4 + j) ^ sw) << 4), 0u, 0u, 0u, 0u);
        continue;
      }
      uint32_t r[32];
      tmem_ld32(tmem_s + lane_addr + c * 32, r);
      tmem_ld_wait();
      float pv[32];
      float l4[4] = {0.f, 0.f, 0.f, 0.f};     // four independent partial sums (one running sum is a 32-deep FADD chain per chunk)
      if (cm == 0xffffffffu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          pv[j] = ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2, -mb));
          l4[j & 3] += pv[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float e = ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2, -mb));
          pv[j] = ((cm >> j) & 1u) ? e : 0.f;
          l4[j & 3] += pv[j];
        }
      }
      lsum += (l4[0] + l4[1]) + (l4[2] + l4[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        st_shared_v4(pb + ((((c & 1) * 4 + j) ^ sw) << 4), pack_bf16(pv[8 * j], pv[8 * j + 1]), pack_bf16(pv[8 * j + 2], pv[8 * j + 3]),
                     pack_bf16(pv[8 * j + 4], pv[8 * j + 5]), pack_bf16(pv[8 * j + 6], pv[8 * j + 7]));
    }
    psum[half][row] = lsum;
    ATC2_TR();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    ATC2_TR();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint64_t da = umma_desc_sw128(qs + (j >> 2) * ATC_TILE) + 2 * (j & 3);
          const uint64_t db = umma_desc_sw128_mn(vs + j * 2048);
          umma_bf16(tmem_o, da, db, idesc_o, j != 0 ? 1u : 0u);
        }
        umma_commit(&bar_o);
      }
      __syncwarp();
    }
    const float ltot = psum[0][row] + psum[1][row];
    const float inv = ltot > 0.f ? 1.f / ltot : 0.f;
    const int b = unit_utt(unit, p.n_heads), head = unit - b * p.n_heads;
    ATC2_TR();
    mbar_wait(&bar_o, phase);
    tc_fence_after();
    ATC2_TR();
    {  // out row = O / rowsum: this warp's 32 of the 64 columns, staged in the (dead) Q tile so that the global stores are
       // full 128-byte rows (row-per-thread stores touched 32 lines per instruction: ~1 k cycles of LSU time per unit)
      uint32_t r[32];
      tmem_ld32(tmem_o + lane_addr + half * 32, r);
      tmem_ld_wait();
      const uint32_t orow = qs + row * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        st_shared_v4(orow + (((half * 4 + j) ^ (row & 7)) << 4),
                     pack_bf16(__uint_as_float(r[8 * j]) * inv, __uint_as_float(r[8 * j + 1]) * inv),
                     pack_bf16(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv),
                     pack_bf16(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv),
                     pack_bf16(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv));
    }
    tc_fence_before();
    __syncthreads();     // O tile complete; TMEM and the row-statistic arrays are free again
    {
      __nv_bfloat16* obase = p.out + static_cast<size_t>(b) * p.n_q * p.ldo + head * ATT_DH;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int c = it * 256 + tid, orow_i = c >> 3, ch = c & 7;     // 8 x 16 B per 128-byte row
        const int otok = orow_i & 63, obr = orow_i >> 6;
        if (otok < n_tok && !(p.single && obr)) {
          const uint4 v = lds_u4(qs + orow_i * 128 + ((ch ^ (orow_i & 7)) << 4));
          *reinterpret_cast<uint4*>(obase + static_cast<size_t>(p.single ? otok : 2 * otok + obr) * p.ldo + ch * 8) = v;
        }
      }
    }
    fence_proxy_async();   // the next TMA boxes (async proxy) overwrite the staged tile
    __syncthreads();     // this buffer is free again
    ATC2_TR();
    buf ^= 1;
    phase ^= 1u;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_slot);
}

// ------------------------------------------------------------------------------------------------------
// attention_tc4_kernel — the same contractions with ONE unit in flight per CTA and FOUR CTAs per SM.
//
// attention_tc2_kernel's timeline (tools/att_trace.py) is a serial chain per unit — operands land, S = Q K^T, row max,
// exp + P, P V, normalise + store: ~4.7 k cycles — and at cfg2 (512 units) its 296 persistent CTAs run that chain twice
// behind a ~2 k-cycle issue prologue.  Here a CTA is 4 warps (thread = query row = TMEM lane: no cross-warp row
// statistics, two block barriers fewer), one operand buffer (48 KB) and 128 TMEM columns (the O accumulator
// overwrites S, which is dead once P is in shared memory), so four CTAs share an SM and all 512 units of cfg2 run in
// one wave: the chain is paid once and the four CTAs fill each other's bubbles.
// ------------------------------------------------------------------------------------------------------
constexpr int ATC4_SMEM_BYTES = ATC_BUF_BYTES + 1024;

__global__ void __launch_bounds__(128, 4) attention_tc4_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                               const __grid_constant__ CUtensorMap tmT,
                                                               const __grid_constant__ CUtensorMap tmP,
                                                               const __grid_constant__ CUtensorMap tmN,
                                                               const AttnTcParams p) {
  extern __shared__ uint8_t atc_smem_raw[];
  __shared__ __align__(8) uint64_t bar_full, bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t colmask[2][4];        // [branch][32-key chunk]
  const uint32_t smem_base = (smem_u32(atc_smem_raw) + 1023u) & ~1023u;
  const uint32_t qs = smem_base, ks = qs + ATC_TILE, vs = ks + ATC_TILE;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp index provably uniform
  const int row = tid;                                 // query row = TMEM lane
  const int br = row >> 6;                             // rows 0..63 conditional, 64..127 unconditional (warp-uniform)
  const int n_tok = p.n_style;
#ifdef STZ_TRACE
  long long* tr = (g_att_trace != nullptr && tid == 0 && blockIdx.x < 296) ? g_att_trace + blockIdx.x * 32 : nullptr;
#else
  constexpr long long* tr = nullptr;
#endif
  int tri = 0;
#define ATC4_TR() do { if (tr != nullptr && tri < 32) tr[tri++] = clock64(); } while (0)
  ATC4_TR();

  if (tid == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmT);
    if (!p.self) { prefetch_tmap(&tmP); prefetch_tmap(&tmN); }
    mbar_init(&bar_full, 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<128>(&tmem_slot);
  // V rows that no TMA box covers must be finite (p = 0 times NaN would poison O): see attention_tc2_kernel
  if (!p.self || p.single) {
    for (uint32_t o = tid * 16; o < ATC_TILE; o += 128 * 16) st_shared_v4(vs + o, 0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = tmem_slot, tmem_o = tmem_slot;        // O (64 columns) overwrites S (128 columns)
  ATC4_TR();
  const uint32_t nbox = p.single ? 1u : 2u;
  const uint32_t tx_bytes = p.self ? 3u * nbox * 8192u : nbox * 8192u + 2u * 128u * static_cast<uint32_t>(p.T + p.P + 1);
  // parts: 1 = Q (written by the previous kernel), 2 = K / V.  One elected lane per warp issues that warp's boxes; warp 0
  // arms the barrier with the unit's total bytes (boxes of other warps may land first: the transaction count is signed).
  auto produce_tma = [&](int b, int head, int parts) {
    const uint32_t bar = smem_u32(&bar_full);
    const int t0 = b * n_tok, hc = head * ATT_DH;
    if (!elect_one()) return;
    if (parts & 2) {
      if (warp == 0) mbar_expect_tx(&bar_full, tx_bytes);
      if (p.self) {
        if (p.box2 == 2) {         // Q | K | V tiles in one 4-D box (tmT: column, token, branch, part)
          if (warp == 0) tma_load_4d_u32(qs, &tmT, bar, hc, t0, 0, 0);
        } else if (p.box2) {
          if (warp == 0) tma_load_3d_u32(ks, &tmQ, bar, p.col_k + hc, t0, 0);
          else if (warp == 1) tma_load_3d_u32(vs, &tmQ, bar, p.col_v + hc, t0, 0);
        } else {
          if (warp == 0) {
            tma_load_3d_u32(ks, &tmQ, bar, p.col_k + hc, 0, t0);
            if (!p.single) tma_load_3d_u32(ks + 8192, &tmQ, bar, p.col_k + hc, 1, t0);
          } else if (warp == 1) {
            tma_load_3d_u32(vs, &tmQ, bar, p.col_v + hc, 0, t0);
            if (!p.single) tma_load_3d_u32(vs + 8192, &tmQ, bar, p.col_v + hc, 1, t0);
          }
        }
      } else {
        const uint32_t o1 = p.T8 * 128, o2 = (p.T8 + p.P8) * 128;
        if (warp == 0) {
          tma_load_2d_u32(ks, &tmT, bar, p.col_k + hc, b * p.T);
          tma_load_2d_u32(vs, &tmT, bar, p.col_v + hc, b * p.T);
        } else if (warp == 1) {
          tma_load_2d_u32(ks + o1, &tmP, bar, p.col_k + hc, b * p.P);
          tma_load_2d_u32(vs + o1, &tmP, bar, p.col_v + hc, b * p.P);
        } else if (warp == 2) {
          tma_load_2d_u32(ks + o2, &tmN, bar, p.col_k + hc, 0);
          tma_load_2d_u32(vs + o2, &tmN, bar, p.col_v + hc, 0);
        }
      }
    }
    if (parts & 1) {
      if (warp == 3) {
        if (p.box2 == 2 && p.self) {
        } else if (p.box2) {
          tma_load_3d_u32(qs, &tmQ, bar, hc, t0, 0);
        } else {
          tma_load_3d_u32(qs, &tmQ, bar, hc, 0, t0);
          if (!p.single) tma_load_3d_u32(qs + 8192, &tmQ, bar, hc, 1, t0);
        }
      }
    }
  };
  auto unit_bh = [&](int unit, int& b, int& head) {
    b = unit_utt(unit, p.n_heads);
    head = unit - b * p.n_heads;
  };
  // The cross-attention K / V of a call are constants of the evaluation loop: the first unit's boxes are requested
  // BEFORE griddepcontrol.wait and land while the previous kernel drains.
  const bool kv_early = !p.self && static_cast<int>(blockIdx.x) < p.n_units;
  if (kv_early) {
    int b, head;
    unit_bh(blockIdx.x, b, head);
    produce_tma(b, head, 2);
  }
  pdl_sync();
  ATC4_TR();

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64) | (1u << 16);   // B (= V) is MN-major
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  uint32_t phase = 0;
  for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
    int b, head;
    unit_bh(unit, b, head);
    produce_tma(b, head, (kv_early && unit == static_cast<int>(blockIdx.x)) ? 1 : 3);
    {   // visibility of key row `tid`: bit 0 = conditional queries, bit 1 = unconditional queries
      uint8_t vis = 0;
      if (p.self) {
        vis = (tid & 63) < n_tok ? static_cast<uint8_t>(1u << (tid >> 6)) : 0;
      } else if (tid < p.T) {
        vis = (p.tmask == nullptr || p.tmask[static_cast<size_t>(b) * p.T + tid] != 0) ? 3 : 0;
      } else if (tid >= p.T8 && tid < p.T8 + p.P) {
        vis = (p.pmask == nullptr || p.pmask[static_cast<size_t>(b) * p.P + (tid - p.T8)] != 0) ? 1 : 0;
      } else if (tid == p.T8 + p.P8) {
        vis = 2;
      }
      // warp w holds the visibility of key chunk w
      const uint32_t m0 = __ballot_sync(0xffffffffu, (vis & 1) != 0), m1 = __ballot_sync(0xffffffffu, (vis & 2) != 0);
      if (lane == 0) { colmask[0][warp] = m0; colmask[1][warp] = m1; }
    }
    ATC4_TR();
    mbar_wait(&bar_full, phase);
    __syncthreads();              // colmask visible (and every thread past the barrier wait)
    ATC4_TR();
    if (warp == 0) {
      tc_fence_after();
      const uint64_t da = umma_desc_sw128(qs), db = umma_desc_sw128(ks);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_s, da + 2 * k, db + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&bar_s);
      }
      __syncwarp();
    }
    uint32_t cm[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) cm[c] = colmask[br][c];
    mbar_wait(&bar_s, phase);
    tc_fence_after();
    ATC4_TR();

    // ---- pass 1: row max over the visible keys
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (cm[c] == 0) continue;   // warp-uniform
      uint32_t r[32];
      tmem_ld32(tmem_s + lane_addr + c * 32, r);
      tmem_ld_wait();
      if (cm[c] == 0xffffffffu) {   // four independent chains (a single running max is a 16-deep dependent FMNMX chain per chunk)
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 32; j += 8)
#pragma unroll
          for (int i = 0; i < 4; ++i) m4[i] = fmaxf(m4[i], fmaxf(__uint_as_float(r[j + 2 * i]), __uint_as_float(r[j + 2 * i + 1])));
        mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
      } else {   // partially visible chunk: mask, so that the result never depends on a neighbouring utterance's rows
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, ((cm[c] >> j) & 1u) ? __uint_as_float(r[j]) : -INFINITY);
      }
    }
    const float mb = (mx == -INFINITY ? 0.f : mx) * p.scale_log2;
    ATC4_TR();

    // ---- pass 2: P = exp2(S * scale - max), masked, bf16, into the dead Q | K tiles (K-major, swizzled)
    float lsum = 0.f;
    const uint32_t prow = qs + row * 128, sw = row & 7;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t pb = prow + (c >> 1) * ATC_TILE;
      if (cm[c] == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) st_shared_v4(pb + ((((c & 1) * 4 + j) ^ sw) << 4), 0u, 0u, 0u, 0u);
        continue;
      }
      uint32_t r[32];
      tmem_ld32(tmem_s + lane_addr + c * 32, r);
      tmem_ld_wait();
      float pv[32];
      float l4[4] = {0.f, 0.f, 0.f, 0.f};     // four independent partial sums (one running sum is a 32-deep FADD chain per chunk)
      if (cm[c] == 0xffffffffu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          pv[j] = ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2, -mb));
          l4[j & 3] += pv[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float e = ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2, -mb));
          pv[j] = ((cm[c] >> j) & 1u) ? e : 0.f;
          l4[j & 3] += pv[j];
        }
      }
      lsum += (l4[0] + l4[1]) + (l4[2] + l4[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        st_shared_v4(pb + ((((c & 1) * 4 + j) ^ sw) << 4), pack_bf16(pv[8 * j], pv[8 * j + 1]), pack_bf16(pv[8 * j + 2], pv[8 * j + 3]),
                     pack_bf16(pv[8 * j + 4], pv[8 * j + 5]), pack_bf16(pv[8 * j + 6], pv[8 * j + 7]));
    }
    ATC4_TR();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();              // P complete, every thread done reading S
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint64_t da = umma_desc_sw128(qs + (j >> 2) * ATC_TILE) + 2 * (j & 3);
          const uint64_t db = umma_desc_sw128_mn(vs + j * 2048);
          umma_bf16(tmem_o, da, db, idesc_o, j != 0 ? 1u : 0u);
        }
        umma_commit(&bar_o);
      }
      __syncwarp();
    }
    const float inv = lsum > 0.f ? 1.f / lsum : 0.f;
    ATC4_TR();
    mbar_wait(&bar_o, phase);
    tc_fence_after();
    ATC4_TR();
    {  // out row = O / rowsum, staged in the (dead) Q tile so that the global stores are full 128-byte rows
      const uint32_t orow = qs + row * 128;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t r[32];
        tmem_ld32(tmem_o + lane_addr + h * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(orow + (((h * 4 + j) ^ (row & 7)) << 4),
                       pack_bf16(__uint_as_float(r[8 * j]) * inv, __uint_as_float(r[8 * j + 1]) * inv),
                       pack_bf16(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv),
                       pack_bf16(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv),
                       pack_bf16(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv));
      }
    }
    tc_fence_before();
    __syncthreads();     // O tile complete; TMEM is free again
    {
      __nv_bfloat16* obase = p.out + static_cast<size_t>(b) * p.n_q * p.ldo + head * ATT_DH;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int c = it * 128 + tid, orow_i = c >> 3, ch = c & 7;     // 8 x 16 B per 128-byte row
        const int otok = orow_i & 63, obr = orow_i >> 6;
        if (otok < n_tok && !(p.single && obr)) {
          const uint4 v = lds_u4(qs + orow_i * 128 + ((ch ^ (orow_i & 7)) << 4));
          *reinterpret_cast<uint4*>(obase + static_cast<size_t>(p.single ? otok : 2 * otok + obr) * p.ldo + ch * 8) = v;
        }
      }
    }
    fence_proxy_async();   // the next TMA boxes (async proxy) overwrite the staged tile
    __syncthreads();       // the buffer is free again
    ATC4_TR();
    phase ^= 1u;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem_slot);
}

// ------------------------------------------------------------------------------------------------------
// attention_tcs_kernel — streaming form of attention_tc2_kernel for cross-attention over long text
// (T8 + P8 + 1 > 128 keys, e.g. BASELINE cfg4: 512 text tokens + 50 prompt tokens + null).
//
// The key sequence is visited in blocks of 128 rows: ceil(T / 128) text blocks, then one block holding the prompt
// keys (rows 0..P-1) and the null-prompt key (row P8).  Per block: S = Q K_j^T (tcgen05), thread-per-row masked
// softmax with a running row max / row sum, P_j -> shared memory (the first 64 keys over the dead K_j tile),
// O (+)= P_j V_j (tcgen05, V MN-major).  When a block raises a row's running max the O accumulator of that row is
// rescaled in TMEM (tcgen05.ld -> x alpha -> tcgen05.st); warps skip the rescale when no row of theirs moved.
// K_j / V_j are double-buffered (block j + 1 is fetched while block j is processed, and S_{j+1} is issued right behind
// P_j V_j); Q stays resident for the unit.
// Same work split as attention_tc2_kernel: warps w and w + 4 share 32 query rows and split the four 32-key chunks
// of a block by parity and the 64 output columns in halves.
// ------------------------------------------------------------------------------------------------------
constexpr int ATS_SMEM_BYTES = ATC_TILE /*Q*/ + 2 * 2 * ATC_TILE /*2 x (K, V)*/ + ATC_TILE /*P keys 64..127*/ + 1024;

__global__ void __launch_bounds__(256, 2) attention_tcs_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                               const __grid_constant__ CUtensorMap tmT,
                                                               const __grid_constant__ CUtensorMap tmP,
                                                               const __grid_constant__ CUtensorMap tmN,
                                                               const AttnTcParams p) {
  extern __shared__ uint8_t atc_smem_raw[];
  __shared__ __align__(8) uint64_t bar_q, bar_kv[2], bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t colmask[2][4];
  __shared__ float pmax[2][128], psum[2][128];
  const uint32_t smem_base = (smem_u32(atc_smem_raw) + 1023u) & ~1023u;
  const uint32_t qs = smem_base, kv0 = smem_base + ATC_TILE, p1s = smem_base + 5 * ATC_TILE;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp index provably uniform
  const int q4 = warp & 3, half = warp >> 2;
  const int row = q4 * 32 + lane;
  const int br = row >> 6, tok = row & 63;
  const int n_tok = p.n_style;
  const int nbt = (p.T + 127) >> 7;                        // text blocks; block id nbt = the prompt + null block

  if (tid == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmT); prefetch_tmap(&tmP); prefetch_tmap(&tmN);
    mbar_init(&bar_q, 1);
    mbar_init(&bar_kv[0], 1);
    mbar_init(&bar_kv[1], 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(&tmem_slot);
  for (uint32_t o = tid * 16; o < 6 * ATC_TILE; o += 256 * 16) st_shared_v4(smem_base + o, 0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = tmem_slot, tmem_o = tmem_slot + 128;
  pdl_sync();

  // jj = position in this unit's block sequence (selects the buffer), j = block id (text block, or nbt = prompt + null)
  auto load_block = [&](int b, int head, int jj, int j) {     // tid == 0
    const int buf = jj & 1;
    const uint32_t ks = kv0 + buf * 2 * ATC_TILE, vs = ks + ATC_TILE, bar = smem_u32(&bar_kv[buf]);
    const int hc = head * ATT_DH;
    if (j < nbt) {
      mbar_expect_tx(&bar_kv[buf], 2u * ATC_TILE);
      tma_load_2d_u32(ks, &tmT, bar, p.col_k + hc, b * p.T + j * 128);
      tma_load_2d_u32(vs, &tmT, bar, p.col_v + hc, b * p.T + j * 128);
    } else {
      mbar_expect_tx(&bar_kv[buf], 2u * 128u * static_cast<uint32_t>(p.P + 1));
      tma_load_2d_u32(ks, &tmP, bar, p.col_k + hc, b * p.P);
      tma_load_2d_u32(vs, &tmP, bar, p.col_v + hc, b * p.P);
      tma_load_2d_u32(ks + p.P8 * 128, &tmN, bar, p.col_k + hc, 0);
      tma_load_2d_u32(vs + p.P8 * 128, &tmN, bar, p.col_v + hc, 0);
    }
  };

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64) | (1u << 16);
  const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
  uint32_t ph_q = 0, ph_kv0 = 0, ph_kv1 = 0, ph_s = 0, ph_o = 0;

  // tid 0 only: S_j = Q K_j^T as soon as block j has landed (issued right behind the previous block's P V, so its latency
  // hides under that block's tail instead of heading the next block's dependency chain)
  auto issue_s = [&](int jj) {
    const int buf = jj & 1;
    if (jj == 0) { mbar_wait(&bar_q, ph_q); ph_q ^= 1u; }
    if (buf == 0) { mbar_wait(&bar_kv[0], ph_kv0); ph_kv0 ^= 1u; } else { mbar_wait(&bar_kv[1], ph_kv1); ph_kv1 ^= 1u; }
    tc_fence_after();
    const uint64_t da = umma_desc_sw128(qs), db = umma_desc_sw128(kv0 + buf * 2 * ATC_TILE);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem_s, da + 2 * k, db + 2 * k, idesc_s, k != 0 ? 1u : 0u);
    umma_commit(&bar_s);
  };

  for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
    const int b = unit_utt(unit, p.n_heads), head = unit - b * p.n_heads;
    // text masks are prefix masks: a text block whose first key is padding is padding throughout and is skipped
    int nbv = nbt;
    if (p.tmask != nullptr) {
      nbv = 1;
      for (int j = 1; j < nbt; ++j) nbv += p.tmask[static_cast<size_t>(b) * p.T + j * 128] != 0 ? 1 : 0;
    }
    const int nblk = nbv + 1;                  // valid text blocks, then the prompt + null block (id nbt)
    if (warp == 0 && elect_one()) {     // one elected lane of warp 0 issues (elect.sync region: no per-instruction waterfall)
      mbar_expect_tx(&bar_q, p.single ? 8192u : 2u * 8192u);
      if (p.box2) {
        tma_load_3d_u32(qs, &tmQ, smem_u32(&bar_q), head * ATT_DH, b * n_tok, 0);
      } else {
        tma_load_3d_u32(qs, &tmQ, smem_u32(&bar_q), head * ATT_DH, 0, b * n_tok);
        if (!p.single) tma_load_3d_u32(qs + 8192, &tmQ, smem_u32(&bar_q), head * ATT_DH, 1, b * n_tok);
      }
      load_block(b, head, 0, 0);
      load_block(b, head, 1, nblk > 2 ? 1 : nbt);       // nblk >= 2 always
      issue_s(0);
    }
    float m_run = -INFINITY, l_run = 0.f;
    for (int jj = 0; jj < nblk; ++jj) {
      const int j = jj < nbv ? jj : nbt;         // block id
      const int buf = jj & 1;
      const uint32_t ks = kv0 + buf * 2 * ATC_TILE, vs = ks + ATC_TILE;
      if (warp < 4) {   // visibility of key row warp * 32 + lane of this block -> per-branch 32-key chunk masks
        const int kr = warp * 32 + lane;
        uint32_t vis = 0;
        if (j < nbt) {
          const int kt = j * 128 + kr;
          vis = (kt < p.T && (p.tmask == nullptr || p.tmask[static_cast<size_t>(b) * p.T + kt] != 0)) ? 3 : 0;
        } else if (kr < p.P) {
          vis = (p.pmask == nullptr || p.pmask[static_cast<size_t>(b) * p.P + kr] != 0) ? 1 : 0;
        } else if (kr == p.P8) {
          vis = 2;
        }
        const uint32_t m0 = __ballot_sync(0xffffffffu, (vis & 1) != 0), m1 = __ballot_sync(0xffffffffu, (vis & 2) != 0);
        if (lane == 0) { colmask[0][warp] = m0; colmask[1][warp] = m1; }
      }
      __syncthreads();            // colmask visible
      const uint32_t cm0 = colmask[br][half], cm1 = colmask[br][half + 2];
      mbar_wait(&bar_s, ph_s); ph_s ^= 1u;
      // the previous block's P V has completed: its P tiles may be overwritten, O may be rescaled, and its K / V buffer
      // may be refilled with block j + 1
      if (jj > 0) {
        mbar_wait(&bar_o, ph_o); ph_o ^= 1u;
        if (warp == 0 && jj + 1 < nblk && elect_one()) load_block(b, head, jj + 1, jj + 1 < nbv ? jj + 1 : nbt);
      }
      tc_fence_after();

      // ---- pass 1: block row max over the visible keys of this warp's chunks
      float mx = -INFINITY;
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const uint32_t cm = ci ? cm1 : cm0;
        if (cm == 0) continue;
        uint32_t r[32];
        tmem_ld32(tmem_s + lane_addr + (half + 2 * ci) * 32, r);
        tmem_ld_wait();
        if (cm == 0xffffffffu) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, ((cm >> i) & 1u) ? __uint_as_float(r[i]) : -INFINITY);
        }
      }
      pmax[half][row] = mx;
      __syncthreads();
      const float m_new = fmaxf(m_run, fmaxf(pmax[0][row], pmax[1][row]));
      const float mb = (m_new == -INFINITY ? 0.f : m_new) * p.scale_log2;
      const float alpha = (m_run == -INFINITY) ? (m_new == -INFINITY ? 1.f : 0.f) : ex2_approx((m_run - m_new) * p.scale_log2);
      m_run = m_new;

      // ---- pass 2: P_j = exp2(S * scale - max), masked, bf16 (keys 0..63 over the dead K_j tile, keys 64..127 in p1s)
      float lsum = 0.f;
      const uint32_t sw = row & 7;
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int c = half + 2 * ci;
        const uint32_t cm = ci ? cm1 : cm0;
        const uint32_t pb = ((c >> 1) ? p1s : ks) + row * 128;
        if (cm == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) st_shared_v4(pb + ((((c & 1) * 4 + i) ^ sw) << 4), 0u, 0u, 0u, 0u);
          continue;
        }
        uint32_t r[32];
        tmem_ld32(tmem_s + lane_addr + c * 32, r);
        tmem_ld_wait();
        float pv[32];
        if (cm == 0xffffffffu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) { pv[i] = ex2_approx(fmaf(__uint_as_float(r[i]), p.scale_log2, -mb)); lsum += pv[i]; }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = ex2_approx(fmaf(__uint_as_float(r[i]), p.scale_log2, -mb));
            pv[i] = ((cm >> i) & 1u) ? e : 0.f;
            lsum += pv[i];
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          st_shared_v4(pb + ((((c & 1) * 4 + i) ^ sw) << 4), pack_bf16(pv[8 * i], pv[8 * i + 1]), pack_bf16(pv[8 * i + 2], pv[8 * i + 3]),
                       pack_bf16(pv[8 * i + 4], pv[8 * i + 5]), pack_bf16(pv[8 * i + 6], pv[8 * i + 7]));
      }
      psum[half][row] = lsum;
      // ---- rescale this warp's 32 columns of O where the running max moved
      if (jj > 0) {
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
          uint32_t r[32];
          tmem_ld32(tmem_o + lane_addr + half * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st32(tmem_o + lane_addr + half * 32, r);
          tmem_st_wait();
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      l_run = l_run * alpha + (psum[0][row] + psum[1][row]);
      if (warp == 0 && elect_one()) {
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint64_t da = umma_desc_sw128((i >> 2) ? p1s : ks) + 2 * (i & 3);
          const uint64_t db = umma_desc_sw128_mn(vs + i * 2048);
          umma_bf16(tmem_o, da, db, idesc_o, (jj | i) != 0 ? 1u : 0u);
        }
        umma_commit(&bar_o);
        if (jj + 1 < nblk) issue_s(jj + 1);      // S is free: every warp finished reading S_j before the barrier above
      }
    }
    // ---- out row = O / rowsum
    mbar_wait(&bar_o, ph_o); ph_o ^= 1u;
    tc_fence_after();
    {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      __nv_bfloat16* op = p.out + (static_cast<size_t>(b) * p.n_q + (p.single ? tok : 2 * tok + br)) * p.ldo + head * ATT_DH + half * 32;
      uint32_t r[32];
      tmem_ld32(tmem_o + lane_addr + half * 32, r);
      tmem_ld_wait();
      if (tok < n_tok && !(p.single && br)) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(r[8 * i]) * inv, __uint_as_float(r[8 * i + 1]) * inv);
          u.y = pack_bf16(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv);
          u.z = pack_bf16(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv);
          u.w = pack_bf16(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(op + i * 8) = u;
        }
      }
    }
    tc_fence_before();
    __syncthreads();     // TMEM, Q tile and statistics are free for the next unit
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_slot);
}

}  // namespace stz
