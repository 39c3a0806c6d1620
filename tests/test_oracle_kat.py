"""Analytic known-answer tests that pin the oracle.  The reference publishes no
tests or golden vectors (/root/reference/README.md:11-16), so these closed-form identities are what
keep the oracle from being merely self-consistent."""
import math

import pytest
import torch

import styletts_zs_b200 as stz
from oracle import schedule as S
from oracle.model import OraclePath, sample_loop, SAMPLER_STUDENT, SAMPLER_TEACHER

CFG = stz.DEFAULT


def test_karras_schedule_closed_form():
    for n in (2, 4, 33, 65):
        s = S.karras_sigmas(n, CFG.sigma_min, CFG.sigma_max, CFG.rho)
        assert len(s) == n
        assert s[0] == pytest.approx(CFG.sigma_max, rel=1e-12)
        assert s[-1] == pytest.approx(CFG.sigma_min, rel=1e-9)
        assert all(a > b for a, b in zip(s, s[1:]))
    # rho-law midpoint of a 3-point grid
    s = S.karras_sigmas(3, 1e-4, 3.0, 9.0)
    mid = (0.5 * (3.0 ** (1 / 9) + 1e-4 ** (1 / 9))) ** 9
    assert s[1] == pytest.approx(mid, rel=1e-12)
    assert S.karras_sigmas(1, 1e-4, 3.0, 9.0) == [3.0]
    assert S.student_sigmas(1, CFG) == [3.0, 0.0]
    assert len(S.student_sigmas(4, CFG)) == 5 and S.student_sigmas(4, CFG)[-1] == 0.0
    assert len(S.teacher_sigmas(32, CFG)) == 33


def test_edm_precondition_identities():
    sd = CFG.sigma_data
    for sigma in (1e-4, 0.01, 0.5, 3.0, 80.0):
        c_skip, c_out, c_in, c_noise = S.edm_precond(sigma, sd)
        assert c_in ** 2 * (sigma ** 2 + sd ** 2) == pytest.approx(1.0, rel=1e-12)
        assert c_out ** 2 == pytest.approx(sigma ** 2 * c_skip, rel=1e-12)
        assert c_noise == pytest.approx(math.log(sigma) / 4)
    c_skip, c_out, _, _ = S.edm_precond(1e-9, sd)
    assert c_skip == pytest.approx(1.0) and c_out == pytest.approx(0.0, abs=1e-8)
    c_skip, c_out, _, _ = S.edm_precond(1e9, sd)
    assert c_skip == pytest.approx(0.0, abs=1e-12) and c_out == pytest.approx(sd, rel=1e-9)


def _gauss_denoiser(x, sigma):
    """Exact posterior mean for data ~ N(0, sigma_data^2): D = x sd^2 / (sd^2 + sigma^2) (i.e. F = 0)."""
    sd2 = CFG.sigma_data ** 2
    return x * sd2 / (sd2 + sigma * sigma)


def test_euler_student_converges_to_pf_ode_solution():
    """PF-ODE under the Gaussian denoiser: x(s) = x(s0) sqrt((sd^2+s^2)/(sd^2+s0^2)); terminal s = 0."""
    sd2, s0 = CFG.sigma_data ** 2, CFG.sigma_max
    noise = torch.randn(1, 3, 4, 8, generator=torch.Generator().manual_seed(0), dtype=torch.float64)
    exact = s0 * noise[0] * math.sqrt(sd2 / (sd2 + s0 * s0))
    errs = []
    for n in (4, 16, 64, 256, 1024):
        x = sample_loop(CFG, _gauss_denoiser, noise, n, SAMPLER_STUDENT)
        errs.append(float((x - exact).abs().max() / exact.abs().max()))
    assert all(a > b for a, b in zip(errs, errs[1:])), errs
    assert errs[-1] < 5e-3, errs
    # one step lands exactly on D(x0, sigma_max)
    x1 = sample_loop(CFG, _gauss_denoiser, noise, 1, SAMPLER_STUDENT)
    assert torch.allclose(x1, _gauss_denoiser(s0 * noise[0], s0), rtol=1e-12, atol=0)


def test_adpm2_step_is_second_order_and_variance_preserving():
    sd2 = CFG.sigma_data ** 2
    # (1) local order: with the ancestral noise zeroed a step integrates the ODE from s to s_down with
    #     the explicit midpoint rule -> local error O(h^3) or better
    s, x = 1.0, torch.tensor(1.0, dtype=torch.float64)
    errs, hs = [], []
    for h in (0.2, 0.1, 0.05):
        sn = s - h
        s_up, s_down, s_mid = S.adpm2_sigmas(s, sn)
        d = (x - _gauss_denoiser(x, s)) / s
        x_mid = x + d * (s_mid - s)
        d_mid = (x_mid - _gauss_denoiser(x_mid, s_mid)) / s_mid
        x_new = x + d_mid * (s_down - s)
        exact = x * math.sqrt((sd2 + s_down ** 2) / (sd2 + s ** 2))
        errs.append(abs(float(x_new - exact)))
        hs.append(s - s_down)
        assert s_up ** 2 + s_down ** 2 == pytest.approx(sn ** 2, rel=1e-12)
    for i in range(2):  # observed local order >= 3 (global order >= 2)
        assert math.log(errs[i] / errs[i + 1]) / math.log(hs[i] / hs[i + 1]) > 2.9, (errs, hs)
    # (2) the full 32-step teacher maps N(0, sd^2 + smax^2) noise to N(0, ~sd^2) samples
    g = torch.Generator().manual_seed(1)
    steps = 32
    n0 = torch.randn(steps + 1, 1, 1000, 200, generator=g, dtype=torch.float64)
    n0[0] *= math.sqrt(sd2 + CFG.sigma_max ** 2) / CFG.sigma_max   # x0 = smax * noise0 ~ N(0, sd^2 + smax^2)
    x = sample_loop(CFG, _gauss_denoiser, n0, steps, SAMPLER_TEACHER)
    assert float(x.std()) == pytest.approx(math.sqrt(sd2 + CFG.sigma_min ** 2), rel=0.02)


def test_fused_affine_coefficients_match_explicit_step():
    """a-6: x' = alpha x + beta F with alpha = 1 + (1 - c_skip) r, beta = -c_out r (what the CUDA
    epilogue evaluates) equals the explicit Euler / ADPM2 updates."""
    g = torch.Generator().manual_seed(2)
    x, F1, F2, nz = (torch.randn(5, 7, generator=g, dtype=torch.float64) for _ in range(4))
    sd = CFG.sigma_data
    s, sn = 1.7, 0.9
    c_skip, c_out, _, _ = S.edm_precond(s, sd)
    D = c_skip * x + c_out * F1
    explicit = x + (x - D) / s * (sn - s)
    r = (sn - s) / s
    assert torch.allclose(explicit, (1 + (1 - c_skip) * r) * x + (-c_out * r) * F1, rtol=1e-12)
    s_up, s_down, s_mid = S.adpm2_sigmas(s, sn)
    x_mid = x + (x - D) / s * (s_mid - s)
    r1 = (s_mid - s) / s
    assert torch.allclose(x_mid, (1 + (1 - c_skip) * r1) * x - c_out * r1 * F1, rtol=1e-12)
    cs2, co2, _, _ = S.edm_precond(s_mid, sd)
    D2 = cs2 * x_mid + co2 * F2
    explicit2 = x + (x_mid - D2) / s_mid * (s_down - s) + s_up * nz
    r2 = (s_down - s) / s_mid
    assert torch.allclose(explicit2, x + (1 - cs2) * r2 * x_mid - co2 * r2 * F2 + s_up * nz, rtol=1e-12)


def test_round_is_half_to_even():
    v = torch.tensor([0.5, 1.5, 2.5, 3.5, -0.5, 12.5])
    assert torch.round(v).tolist() == [0.0, 2.0, 2.0, 4.0, -0.0, 12.0]


# ---------------------------------------------------------------------------- model properties
def _tiny(B=3, T=12, steps=2, sampler=0, seed=7, var_len=None):
    cfg = stz.TINY
    inp = stz.synthetic_inputs(cfg, B, T, steps=steps, sampler=sampler, seed=seed, var_len=var_len)
    return cfg, inp


def test_cfg_scale_semantics(tiny_weights):
    cfg, inp = _tiny(steps=1)
    o = OraclePath(cfg, tiny_weights)
    run = lambda w: o.sample_style(inp["text_emb"], inp["prompt_feats"], 1, w, noise=inp["noise"])
    z0, z1, z2, z3 = run(0.0), run(1.0), run(2.0), run(3.5)
    # one-step output is affine in the guidance scale
    assert torch.allclose(z2, z0 + 2.0 * (z1 - z0), atol=1e-5)
    assert torch.allclose(z3, z0 + 3.5 * (z1 - z0), atol=1e-5)
    # scale 0 == uncond only: independent of the prompt; scale 1 == cond only: depends on it
    other = inp["prompt_feats"].flip(0) * 1.3
    z0b = o.sample_style(inp["text_emb"], other, 1, 0.0, noise=inp["noise"])
    z1b = o.sample_style(inp["text_emb"], other, 1, 1.0, noise=inp["noise"])
    assert torch.allclose(z0, z0b, atol=1e-6)
    assert not torch.allclose(z1, z1b, atol=1e-3)


def test_padding_invariance(tiny_weights):
    cfg, inp = _tiny(B=3, T=12, steps=2, var_len=(3, 12))
    o = OraclePath(cfg, tiny_weights)
    z = o.sample_style(inp["text_emb"], inp["prompt_feats"], 2, 2.0, text_mask=inp["text_mask"], noise=inp["noise"])
    d = o.predict_duration(inp["text_emb"], z, text_mask=inp["text_mask"])
    # extend T with masked garbage
    T2 = 20
    te = torch.cat([inp["text_emb"], 9.0 * torch.randn(3, T2 - 12, cfg.d_text)], 1)
    te[~torch.cat([inp["text_mask"], torch.zeros(3, T2 - 12, dtype=torch.bool)], 1)] = 123.0
    tm = torch.cat([inp["text_mask"], torch.zeros(3, T2 - 12, dtype=torch.bool)], 1)
    z2 = o.sample_style(te, inp["prompt_feats"], 2, 2.0, text_mask=tm, noise=inp["noise"])
    d2 = o.predict_duration(te, z2, text_mask=tm)
    assert torch.allclose(z, z2, atol=2e-5)
    assert torch.equal(d, d2[:, :12]) and bool((d2[:, 12:] == 0).all())
    assert bool((d[~inp["text_mask"]] == 0).all()) and bool((d[inp["text_mask"]] >= 1).all())


def test_batch_invariance_and_permutation(tiny_weights):
    cfg, inp = _tiny(B=4, T=10, steps=2, var_len=(4, 10))
    o = OraclePath(cfg, tiny_weights)
    kw = dict(text_mask=inp["text_mask"], noise=inp["noise"])
    z = o.sample_style(inp["text_emb"], inp["prompt_feats"], 2, 2.0, **kw)
    d = o.predict_duration(inp["text_emb"], z, text_mask=inp["text_mask"])
    perm = torch.tensor([2, 0, 3, 1])
    zp = o.sample_style(inp["text_emb"][perm], inp["prompt_feats"][perm], 2, 2.0, text_mask=inp["text_mask"][perm],
                        noise=inp["noise"][:, perm])
    dp = o.predict_duration(inp["text_emb"][perm], zp, text_mask=inp["text_mask"][perm])
    assert torch.allclose(z[perm], zp, atol=2e-5) and torch.equal(d[perm], dp)
    # a single utterance on its own gives the same answer -> sharding across GPUs is legal
    z1 = o.sample_style(inp["text_emb"][1:2], inp["prompt_feats"][1:2], 2, 2.0, text_mask=inp["text_mask"][1:2],
                        noise=inp["noise"][:, 1:2])
    assert torch.allclose(z[1:2], z1, atol=2e-5)


def test_determinism(tiny_weights):
    cfg, inp = _tiny(steps=3, sampler=1)
    o = OraclePath(cfg, tiny_weights)
    a = o.sample_style(inp["text_emb"], inp["prompt_feats"], 3, 2.0, noise=inp["noise"], sampler="teacher")
    b = o.sample_style(inp["text_emb"], inp["prompt_feats"], 3, 2.0, noise=inp["noise"], sampler="teacher")
    assert torch.equal(a, b)


def test_bilstm_packed_semantics(tiny_weights):
    """The reverse direction starts at each sequence's own last valid token: the padded batch must
    equal per-sequence unpadded runs (and a hand-written LSTM cell recurrence)."""
    cfg, inp = _tiny(B=3, T=9, var_len=(2, 9))
    o = OraclePath(cfg, tiny_weights)
    style = torch.randn(3, cfg.n_style, cfg.d_style, generator=torch.Generator().manual_seed(3))
    d, s = o.predict_duration(inp["text_emb"], style, text_mask=inp["text_mask"], return_presum=True)
    for b in range(3):
        n = int(inp["lens"][b])
        db, sb = o.predict_duration(inp["text_emb"][b:b + 1, :n], style[b:b + 1], return_presum=True)
        assert torch.allclose(s[b, :n], sb[0], atol=1e-4)
        assert torch.equal(d[b, :n], db[0])
    # hand-written cell vs nn.LSTM on one direction
    W = o.W
    lstm = o._lstms()[0]
    x = torch.randn(1, 5, cfg.d_hid + cfg.d_sty_tok, generator=torch.Generator().manual_seed(4))
    ref, _ = lstm(x)
    h = torch.zeros(cfg.h_lstm); c = torch.zeros(cfg.h_lstm); outs = []
    for t in range(5):
        g = W["lstm0.f.w_ih"] @ x[0, t] + W["lstm0.f.b_ih"] + W["lstm0.f.w_hh"] @ h + W["lstm0.f.b_hh"]
        i, f, gg, og = g.chunk(4)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(og) * torch.tanh(c)
        outs.append(h)
    assert torch.allclose(torch.stack(outs), ref[0, :, :cfg.h_lstm], atol=1e-5)


def test_random_init_is_non_degenerate(default_weights):
    cfg = stz.DEFAULT
    inp = stz.synthetic_inputs(cfg, 1, 24, steps=1)
    o = OraclePath(cfg, default_weights)
    z = o.sample_style(inp["text_emb"], inp["prompt_feats"], 1, 2.0, noise=inp["noise"])
    assert 0.1 < float(z.std()) < 10.0
    zc = o.sample_style(inp["text_emb"], inp["prompt_feats"], 1, 1.0, noise=inp["noise"])
    assert float((z - zc).abs().max()) > 1e-2            # guidance does something
    d = o.predict_duration(inp["text_emb"], z)
    assert int(d.min()) >= 1 and int(d.max()) > int(d.min()) and int(d.max()) <= cfg.max_dur


def test_length_regulator_known_answer(tiny_weights):
    """repeat_interleave semantics, zero-duration tokens, truncation at max_frames, zero padding."""
    o = OraclePath(stz.TINY, tiny_weights)
    f = torch.arange(2 * 4 * 4, dtype=torch.float32).view(2, 4, 4)
    d = torch.tensor([[2, 0, 1, 3], [1, 1, 0, 0]], dtype=torch.int32)
    fr, ln, tk = o.regulate_length(f, d, max_frames=5, return_tokens=True)
    assert ln.tolist() == [5, 2]                                  # 6 frames truncated to 5
    assert tk.tolist() == [[0, 0, 2, 3, 3], [0, 1, -1, -1, -1]]
    assert torch.equal(fr[0, 2], f[0, 2]) and torch.equal(fr[1, 2:], torch.zeros(3, 4))


# ---------------------------------------------------------------------------------------------
# counter-based noise: oracle/philox.py
# ---------------------------------------------------------------------------------------------
def test_philox4x32_10_random123_known_answers():
    """Random123's published kat_vectors for philox4x32, 10 rounds."""
    import numpy as np
    from oracle import philox as PH
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        got = PH.philox4x32_10([np.array([c], dtype=np.uint32) for c in ctr], [np.uint32(k) for k in key])
        assert tuple(int(g[0]) for g in got) == want


def test_philox_normal_is_standard_normal_and_accurate():
    import numpy as np
    from scipy import stats
    from oracle import philox as PH
    z = PH.normal_noise(1234, 0, 2, 8, 50 * 512).astype(np.float64).ravel()
    assert z.dtype == np.float64 and np.isfinite(z).all()
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
    assert abs(stats.skew(z)) < 0.02 and abs(stats.kurtosis(z)) < 0.03
    assert stats.kstest(z[:200000], "norm").pvalue > 1e-3
    # the fp32 polynomial log / sincos against libm in fp64
    u = np.random.default_rng(0).random(100000).astype(np.float32)
    u = u[u > 0]
    assert np.abs(PH._log_u(u) - np.log(u.astype(np.float64))).max() < 2e-6
    k = np.random.default_rng(1).integers(0, 1 << 23, 100000).astype(np.uint32)
    s, c = PH._sincos_turn(k)
    th = 2 * np.pi * (k.astype(np.float64) + 0.5) / 2 ** 23
    assert np.abs(s - np.sin(th)).max() < 3e-7 and np.abs(c - np.cos(th)).max() < 3e-7


def test_philox_noise_is_a_function_of_the_global_utterance_only():
    """Batch / shard invariance: utterance u gets the same noise whichever batch it is drawn in."""
    import numpy as np
    from oracle import philox as PH
    a = PH.normal_noise(7, 0, 3, 6, 2048)
    assert np.array_equal(a[:, 2:5], PH.normal_noise(7, 2, 3, 3, 2048))
    assert np.array_equal(a[1:2], PH.normal_noise(7, 0, 2, 6, 2048)[1:2])
    assert not np.array_equal(a, PH.normal_noise(8, 0, 3, 6, 2048))
    big = PH.normal_noise(7, (1 << 32) + 5, 1, 1, 2048)          # the high utterance word is part of the counter
    assert not np.array_equal(big, PH.normal_noise(7, 5, 1, 1, 2048))


def test_philox_normal_frozen_bit_patterns():
    """Frozen outputs of the generator definition (counter layout, uniform mapping, polynomial Box-Muller): any change to
    oracle/philox.py that alters a bit fails here; the CUDA kernel is held to the same bits in tests/test_gpu_parity.py."""
    import numpy as np
    from oracle import philox as PH
    a = PH.normal_noise(1234, 0, 1, 1, 16)[0, 0].view(np.uint32)
    assert [int(v) for v in a] == [0x3f9d54f8, 0xbfcee928, 0x3eede2d1, 0xbfc7bfcb, 0x3f3e4d4e, 0x3f2211b4, 0x3e40d63d,
                                   0x3dd437c0, 0x3fd46003, 0xbe26305f, 0xbf280d37, 0xbe18edd8, 0x3f07201d, 0x3f22540d,
                                   0x3f87d9a0, 0xbf30e697]
    b = PH.normal_noise(2 ** 63 + 11, (1 << 32) + 3, 3, 2, 8)[2, 1].view(np.uint32)
    assert [int(v) for v in b] == [0x3f177db5, 0xbf19577d, 0xbf812152, 0xbe716d18, 0x3f884dc3, 0xbf1b9418, 0xbe38c012,
                                   0x3fd3134b]
