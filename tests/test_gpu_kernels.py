"""Kernel-level parity: every fused epilogue of the product GEMM kernels (gemm2_kernel's GELU / gated residual reduce-add /
positional / sampler epilogues, gemmln3_kernel's residual + AdaLN passes) against a plain PyTorch fp32 restatement of the
same op on the same bf16 operands, through the C ABI's stz_op_* entry points.  Tolerances are stated per test: the GEMM
itself is exact to fp32 summation order (2e-5); bf16 outputs add one rounding (2^-8 relative)."""
import pytest
import torch

import styletts_zs_b200 as stz

pytestmark = pytest.mark.gpu
CFG = stz.DEFAULT
K_STYLE = CFG.n_style
RPU = 2 * K_STYLE                      # rows per utterance in the R layout


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-12))


@pytest.fixture(scope="module")
def path():
    p = stz.StyleTTSZSPath(CFG, stz.init_weights(CFG, 0), device=0)
    yield p
    p.close()


def _operands(M, N, K, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = 0.5 * torch.randn(N, device="cuda", generator=g)
    return A, W, b, g


def _seq_of_rows(M):
    r = torch.arange(M, device="cuda")
    return (r // RPU) * 2 + (r & 1)


def _n_seq(M):
    return 2 * ((M + RPU - 1) // RPU)


@pytest.mark.parametrize("M,N,K", [(6400, 1536, 512), (300, 512, 512), (3200, 8192, 512)])
def test_epilogue_bf16(path, M, N, K):
    A, W, b, _ = _operands(M, N, K, 1)
    ref = A.float() @ W.float().t() + b
    got = path.op_gemm_epi(A, W, b, 2)
    assert got.dtype == torch.bfloat16 and rel(got, ref) < 5e-3          # one bf16 rounding of the output (2^-8)


@pytest.mark.parametrize("M,N,K", [(6400, 2048, 512), (257, 2048, 512)])
def test_epilogue_gelu_fp16x2(path, M, N, K):
    """FFN1: gelu_tanh in packed fp16 arithmetic (11 significant bits) rounded to bf16."""
    A, W, b, _ = _operands(M, N, K, 2)
    ref = torch.nn.functional.gelu(A.float() @ W.float().t() + b, approximate="tanh")
    got = path.op_gemm_epi(A, W, b, 3)
    assert rel(got, ref) < 6e-3                                          # bf16 output rounding + fp16 tanh.approx


@pytest.mark.parametrize("M,N,K", [(6400, 512, 512), (6400, 512, 2048), (150, 512, 512)])
def test_epilogue_gate_residual_reduce_add(path, M, N, K):
    """h += gate[seq(r)] * (A W^T + b) through TMA reduce-add: h is never read by an SM."""
    A, W, b, g = _operands(M, N, K, 3)
    n_mod, gate_off = 3 * N, N
    mod = torch.randn(_n_seq(M), n_mod, device="cuda", generator=g)
    h0 = torch.randn(M, N, device="cuda", generator=g)
    ref = h0 + mod[_seq_of_rows(M), gate_off:gate_off + N] * (A.float() @ W.float().t() + b)
    h = h0.clone()
    path.op_gemm_epi(A, W, b, 4, out=h, mod=mod, gate_off=gate_off)
    assert rel(h, ref) < 2e-5


@pytest.mark.parametrize("M", [6400, 130])
def test_epilogue_positional(path, M):
    N, K = 512, 1536
    A, W, b, g = _operands(M, N, K, 4)
    pos = torch.randn(K_STYLE, N, device="cuda", generator=g)
    r = torch.arange(M, device="cuda")
    ref = A.float() @ W.float().t() + b + pos[(r // 2) % K_STYLE]
    got = path.op_gemm_epi(A, W, b, 1, pos=pos)
    assert rel(got, ref) < 2e-5


@pytest.mark.parametrize("M,dest,cm,cn", [(6400, 0, 0.0, 0.0), (6400, 1, 0.0, 0.0), (6400, 0, 0.37, 0.8), (102, 0, 0.37, 0.8)])
def test_epilogue_sampler(path, M, dest, cm, cn):
    """a-6: CFG combine + affine sampler update + next input, all inside the output projection's epilogue."""
    N, K = CFG.d_style, 3 * CFG.d_model
    A, W, b, g = _operands(M, N, K, 5)
    x0 = torch.randn(M // 2, N, device="cuda", generator=g)
    xm0 = torch.randn(M // 2, N, device="cuda", generator=g)
    nz = torch.randn(M // 2, N, device="cuda", generator=g)
    cx, cF, cin, w = 0.91, -0.45, 1.7, 2.0
    coef = torch.tensor([cx, cm, cF, cn, cin, w, float(dest), 0.0], device="cuda")
    F = A.float() @ W.float().t() + b
    Fg = F[1::2] + w * (F[0::2] - F[1::2])
    new = cx * x0 + cm * xm0 + cF * Fg + cn * nz
    x, xm = x0.clone(), xm0.clone()
    tap = torch.empty_like(x0)
    xin = path.op_gemm_sampler(A, W, b, x, xm, nz, coef, tap=tap)
    assert rel(tap, Fg) < 3e-5
    if dest == 0:
        assert rel(x, new) < 3e-5 and torch.equal(xm, xm0)
    else:
        assert rel(xm, new) < 3e-5 and torch.equal(x, x0)
    got_state = x if dest == 0 else xm
    y = cin * got_state                                  # the kernel splits ITS state: compare the encoding exactly
    hi = y.bfloat16()
    lo = (y - hi.float()).bfloat16()
    want = torch.cat([hi, lo, hi], 1).repeat_interleave(2, 0)         # both CFG branches get the same input rows
    assert torch.equal(xin, want)


def _ln_mod(hp, mod, seq, shift_off, scale_off):
    N = hp.shape[1]
    ln = torch.nn.functional.layer_norm(hp, (N,), eps=1e-5)
    return ln * (1.0 + mod[seq, scale_off:scale_off + N]) + mod[seq, shift_off:shift_off + N]


@pytest.mark.parametrize("M,K,split3", [(6400, 512, False), (6400, 2048, False), (6400, 2048, True), (333, 512, False), (128, 512, True)])
def test_fused_gemm_residual_adaln(path, M, K, split3):
    """gemmln3_kernel, GLN_RES: h' = h + gate (A W^T + b); u = bf16(LN(h') (1 + scale) + shift).  cfg2's 50 row tiles
    (utterances straddle tiles: 100 rows per utterance vs 128-row tiles), a partial last tile, one tile."""
    N = 512
    A, W, b, g = _operands(M, N, K, 6)
    n_mod = 4 * N
    mod = 0.5 * torch.randn(_n_seq(M), n_mod, device="cuda", generator=g)
    h0 = torch.randn(M, N, device="cuda", generator=g)
    seq = _seq_of_rows(M)
    hp = h0 + mod[seq, 0:N] * (A.float() @ W.float().t() + b)
    u_ref = _ln_mod(hp, mod, seq, N, 2 * N)
    h = h0.clone()
    u = path.op_gemm_ln(A, W, b, h, mod, mode=0, gate_off=0, shift_off=N, scale_off=2 * N, split3=split3)
    assert rel(h, hp) < 2e-5
    if not split3:
        assert rel(u, u_ref) < 5e-3                                     # bf16 output rounding
    else:
        hi, lo, hi2 = u[:, :N], u[:, N:2 * N], u[:, 2 * N:]
        assert torch.equal(hi, hi2)
        assert rel(hi.float() + lo.float(), u_ref) < 5e-5              # split-bf16: ~2^-16


@pytest.mark.parametrize("M", [6400, 200])
def test_fused_gemm_positional_adaln(path, M):
    """gemmln3_kernel, GLN_POS (input projection): h = A W^T + b + pos; u = AdaLN_1 of layer 0."""
    N, K = 512, 1536
    A, W, b, g = _operands(M, N, K, 7)
    mod = 0.5 * torch.randn(_n_seq(M), 2 * N, device="cuda", generator=g)
    pos = torch.randn(K_STYLE, N, device="cuda", generator=g)
    r = torch.arange(M, device="cuda")
    seq = _seq_of_rows(M)
    hp = A.float() @ W.float().t() + b + pos[(r // 2) % K_STYLE]
    u_ref = _ln_mod(hp, mod, seq, 0, N)
    h = torch.full((M, N), float("nan"), device="cuda")                # GLN_POS never reads h
    u = path.op_gemm_ln(A, W, b, h, mod, mode=1, shift_off=0, scale_off=N, pos=pos)
    assert rel(h, hp) < 2e-5 and rel(u, u_ref) < 5e-3


@pytest.mark.parametrize("rows", [8, 24, 64, 72, 88, 96, 104, 120])
@pytest.mark.parametrize("M,K,split3", [(6400, 512, False), (1000, 2048, True)])
def test_fused_gemm_residual_adaln_tile_rows(path, rows, M, K, split3):
    """gemmln3_kernel with fewer than 128 rows per CTA pair (the heuristic picks 88 at cfg2's 6400 rows: 146 SMs busy
    instead of 100): every forced row count gives the same h' / u as the PyTorch restatement."""
    N = 512
    A, W, b, g = _operands(M, N, K, 8)
    mod = 0.5 * torch.randn(_n_seq(M), 4 * N, device="cuda", generator=g)
    h0 = torch.randn(M, N, device="cuda", generator=g)
    seq = _seq_of_rows(M)
    hp = h0 + mod[seq, 0:N] * (A.float() @ W.float().t() + b)
    u_ref = _ln_mod(hp, mod, seq, N, 2 * N)
    h = h0.clone()
    path.set_option("gln_tile_rows", rows)
    try:
        u = path.op_gemm_ln(A, W, b, h, mod, mode=0, gate_off=0, shift_off=N, scale_off=2 * N, split3=split3)
    finally:
        path.set_option("gln_tile_rows", 0)
    assert rel(h, hp) < 2e-5
    if not split3:
        assert rel(u, u_ref) < 5e-3
    else:
        assert torch.equal(u[:, :N], u[:, 2 * N:]) and rel(u[:, :N].float() + u[:, N:2 * N].float(), u_ref) < 5e-5
