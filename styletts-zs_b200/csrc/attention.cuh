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

}  // namespace stz
