"""ORACLE (test infrastructure, not product code): counter-based Gaussian noise, numpy restatement.

SURVEY.md §8(f) rank 4: the sampler's `noise [slices, B, K, Ds]` tensor drawn on the device from a
counter-based generator that a CPU restatement reproduces BIT FOR BIT.  /root/reference holds no
noise code (README.md:11-16), so the generator is pinned here:

  * Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the
    Random123 library's `philox4x32_R(10, ...)`), checked against Random123's published
    known-answer vectors in tests/test_oracle_kat.py;
  * key = (seed & 0xffffffff, seed >> 32); counter = (g, utt & 0xffffffff, slice, utt >> 32) where
    `utt` is the GLOBAL utterance index (first_utterance + b) and g indexes groups of four
    consecutive elements of that utterance's [K*Ds] slice — the noise of an utterance does not
    depend on the batch or the GPU it lands in;
  * the four 32-bit outputs give two Box-Muller pairs (x0,x1) -> (z0,z1), (x2,x3) -> (z2,z3);
  * log / sin / cos are fixed polynomial evaluations made of individually rounded fp32
    multiply / add / divide / sqrt only (no FMA, no libm), so numpy float32 and the CUDA kernel
    (`__fmul_rn`, `__fadd_rn`, `__fdiv_rn`, `__fsqrt_rn`) agree exactly.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.
"""
from __future__ import annotations

import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
f32 = np.float32

# polynomial coefficients, written as the decimal literals the CUDA kernel uses (csrc/philox.cuh)
LN2 = f32(0.693147182)
SQRT2 = f32(1.41421354)
LOG_C = [f32(0.111111112), f32(0.142857149), f32(0.200000003), f32(0.333333343)]  # 1/9, 1/7, 1/5, 1/3
SIN_C = [f32(2.75573188e-06), f32(-1.98412701e-04), f32(8.33333377e-03), f32(-1.66666672e-01)]
COS_C = [f32(-2.75573188e-07), f32(2.48015876e-05), f32(-1.38888892e-03), f32(4.16666679e-02), f32(-0.5)]
ANGLE_STEP = f32(7.49014077e-07)  # (pi / 2) * 2^-21


def philox4x32_10(ctr, key):
    """ctr: 4 uint32 arrays (same shape), key: 2 uint32 arrays or scalars -> 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) for x in ctr]
    k0 = np.asarray(key[0], dtype=np.uint64)
    k1 = np.asarray(key[1], dtype=np.uint64)
    for r in range(10):
        if r:
            k0 = (k0 + np.uint64(_W0)) & _MASK
            k1 = (k1 + np.uint64(_W1)) & _MASK
        p0 = _M0 * c[0]
        p1 = _M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
    return [x.astype(np.uint32) for x in c]


def _log_u(u):
    """ln(u) for fp32 u in (0, 1): exponent split, m in (sqrt2/2, sqrt2], 2 atanh((m-1)/(m+1))."""
    bits = u.view(np.uint32)
    e = (bits >> np.uint32(23)).astype(np.int32) - 127
    m = ((bits & np.uint32(0x007FFFFF)) | np.uint32(0x3F800000)).view(f32)
    big = m > SQRT2
    m = np.where(big, m * f32(0.5), m)
    e = np.where(big, e + 1, e)
    s = (m - f32(1.0)) / (m + f32(1.0))
    s2 = s * s
    p = LOG_C[0]
    for c in LOG_C[1:]:
        p = p * s2 + c
    p = p * s2 + f32(1.0)
    lnm = (f32(2.0) * s) * p
    return e.astype(f32) * LN2 + lnm


def _sincos_turn(k):
    """(sin, cos) of 2 pi (k + 0.5) / 2^23 for 23-bit integers k."""
    q = k >> np.uint32(21)
    j = k & np.uint32(0x1FFFFF)
    swap = j >= np.uint32(1 << 20)
    j = np.where(swap, np.uint32((1 << 21) - 1) - j, j)
    phi = (j.astype(f32) + f32(0.5)) * ANGLE_STEP  # (0, pi/4)
    x2 = phi * phi
    ps = SIN_C[0]
    for c in SIN_C[1:]:
        ps = ps * x2 + c
    sn = phi * (ps * x2 + f32(1.0))
    pc = COS_C[0]
    for c in COS_C[1:]:
        pc = pc * x2 + c
    cs = pc * x2 + f32(1.0)
    sq = np.where(swap, cs, sn)
    cq = np.where(swap, sn, cs)
    s = np.where(q == 0, sq, np.where(q == 1, cq, np.where(q == 2, -sq, -cq)))
    c = np.where(q == 0, cq, np.where(q == 1, -sq, np.where(q == 2, -cq, sq)))
    return s.astype(f32), c.astype(f32)


def box_muller(xa, xb):
    """Two uint32 arrays -> two fp32 N(0,1) arrays (r cos, r sin)."""
    u = ((xa >> np.uint32(9)).astype(f32) + f32(0.5)) * f32(2.0 ** -23)
    r = np.sqrt(f32(-2.0) * _log_u(u))
    s, c = _sincos_turn(xb >> np.uint32(9))
    return r * c, r * s


def normal_noise(seed: int, first_utterance, slices: int, B: int, n_per_utt: int) -> np.ndarray:
    """-> fp32 [slices, B, n_per_utt]; n_per_utt % 4 == 0.  `first_utterance`: int (utterance b is global utterance
    first_utterance + b) or a sequence of B explicit global indices."""
    if n_per_utt % 4:
        raise ValueError("n_per_utt must be a multiple of 4")
    G = n_per_utt // 4
    g = np.arange(G, dtype=np.uint64)[None, None, :]
    if isinstance(first_utterance, (int, np.integer)):
        utt = np.arange(B, dtype=np.uint64) + np.uint64(first_utterance)
    else:
        utt = np.asarray([int(i) for i in first_utterance], dtype=np.uint64)
        if utt.shape != (B,):
            raise ValueError("need one global index per utterance")
    utt = utt[None, :, None]
    sl = np.arange(slices, dtype=np.uint64)[:, None, None]
    shape = (slices, B, G)
    ctr = [np.broadcast_to(g, shape), np.broadcast_to(utt & _MASK, shape), np.broadcast_to(sl, shape),
           np.broadcast_to(utt >> np.uint64(32), shape)]
    key = (np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF))
    x = philox4x32_10(ctr, key)
    z0, z1 = box_muller(x[0], x[1])
    z2, z3 = box_muller(x[2], x[3])
    out = np.stack([z0, z1, z2, z3], axis=-1).astype(f32)
    return out.reshape(slices, B, n_per_utt)
