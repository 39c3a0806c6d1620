"""Hazard evidence of our own (compute-sanitizer is closed on this GPU pool — `gpurun` refuses it; profiles/README.md):
  * memory: the library's workspace arenas are re-planned with a poisoned guard gap behind every internal buffer
    (set_option("guard_bytes")), caller-owned outputs sit inside canary-filled allocations, and after a battery that reaches
    every kernel of the path no guard byte and no canary has changed;
  * races: every asynchronous protocol of the kernels (TMA loads issued before griddepcontrol.wait, mbarrier-paced rings,
    DSMEM exchange of the recurrent state, TMA reduce-add) must give bit-identical results run after run, with programmatic
    dependent launch on and off, captured as a graph or launched eagerly."""
import pytest
import torch

import styletts_zs_b200 as stz

pytestmark = pytest.mark.gpu
CFG = stz.DEFAULT


@pytest.fixture(scope="module")
def weights():
    return stz.init_weights(CFG, 0)


def _battery(p):
    """Reaches every kernel family; returns a list of result tensors."""
    out = []
    inp = stz.synthetic_inputs(CFG, 64, 64, steps=2, seed=1)                                         # fused GEMM + AdaLN, resident attention
    z = p.sample_style(inp["text_emb"], inp["prompt_feats"], 2, 2.0, noise=inp["noise"])
    out += [z, p.predict_duration(inp["text_emb"], z)]
    inp = stz.synthetic_inputs(CFG, 5, 300, steps=2, sampler=stz.SAMPLER_TEACHER, seed=2, var_len=(17, 300))   # streaming attention, masks
    z = p.sample_style(inp["text_emb"], inp["prompt_feats"], 2, 2.0, text_mask=inp["text_mask"], noise=inp["noise"], sampler="teacher")
    out += [z, p.predict_duration(inp["text_emb"], z, text_mask=inp["text_mask"])]
    inp = stz.synthetic_inputs(CFG, 2, 20, P=140, steps=1, seed=3)                                   # mma.sync fallback
    out.append(p.sample_style(inp["text_emb"], inp["prompt_feats"], 1, 2.0, noise=inp["noise"]))
    inp = stz.synthetic_inputs(CFG, 3, 40, steps=2, seed=4, var_len=(10, 40))                        # guided student + seeded noise
    out.append(p.sample_style(inp["text_emb"], inp["prompt_feats"], 2, 2.0, text_mask=inp["text_mask"], seed=5, sampler="guided"))
    style = 0.7 * torch.randn(3, CFG.n_style, CFG.d_style, generator=torch.Generator().manual_seed(1))
    out += list(p.predict_prosody(inp["text_emb"], style, text_mask=inp["text_mask"], max_frames=500))[:3]
    inp = stz.synthetic_inputs(CFG, 130, 24, steps=1, seed=6, var_len=(5, 24))                       # 24 sequences per BiLSTM cluster
    style = 0.7 * torch.randn(130, CFG.n_style, CFG.d_style, generator=torch.Generator().manual_seed(2))
    out.append(p.predict_duration(inp["text_emb"], style, text_mask=inp["text_mask"]))
    a = stz.synthetic_inputs(CFG, 4, 32, steps=2, seed=5)                                            # pipelined host slots
    o0 = p.synthesize_host(a["text_emb"], a["prompt_feats"], 2, 2.0, noise=a["noise"], slot=0)
    o1 = p.synthesize_host(a["text_emb"], a["prompt_feats"], 2, 2.0, seed=7, slot=1)
    p.synthesize_host_wait(0); p.synthesize_host_wait(1)
    out += [o0[0].clone(), o0[1].clone(), o1[0].clone()]
    torch.cuda.synchronize()
    return out


def test_no_kernel_writes_outside_its_buffers(weights):
    p = stz.StyleTTSZSPath(CFG, weights, device=0)
    try:
        p.set_option("guard_bytes", 4096)
        res = _battery(p)
        assert all(bool(torch.isfinite(t.float()).all()) for t in res)
        gaps, bad = p.check_guards()
        assert gaps > 60 and bad == 0, f"{bad} guard bytes overwritten in {gaps} gaps"
        # the same with the fused kernel forced at small batch and graphs off (eager launches take other code paths)
        p.set_option("fuse_ln", 4); p.set_option("use_graph", 0)
        _battery(p)
        assert p.check_guards()[1] == 0
        # caller-owned buffers: outputs inside canary-filled allocations
        inp = stz.synthetic_inputs(CFG, 3, 24, steps=1, seed=9)
        big = torch.full((5, CFG.n_style, CFG.d_style), 1234.5, device="cuda")
        p.sample_style(inp["text_emb"], inp["prompt_feats"], 1, 2.0, noise=inp["noise"], out=big[1:4])
        torch.cuda.synchronize()
        assert bool((big[0] == 1234.5).all()) and bool((big[4] == 1234.5).all()) and bool(torch.isfinite(big[1:4]).all())
        g = torch.Generator(device="cuda").manual_seed(0)
        M, N, K = 333, 512, 512
        A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        W = (torch.randn(N, K, device="cuda", generator=g) / 22).bfloat16()
        bias = torch.randn(N, device="cuda", generator=g)
        mod = torch.randn(8, 3 * N, device="cuda", generator=g)
        for epi, dt in ((0, torch.float32), (2, torch.bfloat16), (3, torch.bfloat16)):
            buf = torch.full((M + 256, N), 7.0, device="cuda", dtype=dt)
            p.op_gemm_epi(A, W, bias, epi, out=buf[128:128 + M])
            torch.cuda.synchronize()
            assert bool((buf[:128] == 7.0).all()) and bool((buf[128 + M:] == 7.0).all()), epi     # TMA stores clip at row M
        hbuf = torch.full((M + 256, N), 7.0, device="cuda")
        hbuf[128:128 + M].normal_(generator=g)
        p.op_gemm_ln(A, W, bias, hbuf[128:128 + M], mod, gate_off=0, shift_off=N, scale_off=2 * N)
        torch.cuda.synchronize()
        assert bool((hbuf[:128] == 7.0).all()) and bool((hbuf[128 + M:] == 7.0).all())
    finally:
        p.close()


def test_results_do_not_depend_on_launch_mode_or_repetition(weights):
    """Bit-identical: run after run (30x at cfg2 size), PDL on vs off, graph vs eager.  A race in the early-issued TMA loads,
    the mbarrier rings or the DSMEM exchange would show as a differing bit sooner or later."""
    p = stz.StyleTTSZSPath(CFG, weights, device=0)
    try:
        ref = _battery(p)
        for _ in range(3):
            for a, b in zip(ref, _battery(p)):
                assert torch.equal(a, b)
        inp = stz.synthetic_inputs(CFG, 64, 64, steps=4, seed=1234)
        dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}
        z0 = p.sample_style(dev["text_emb"], dev["prompt_feats"], 4, 2.0, noise=dev["noise"])
        d0 = p.predict_duration(dev["text_emb"], z0)
        for _ in range(30):
            z = p.sample_style(dev["text_emb"], dev["prompt_feats"], 4, 2.0, noise=dev["noise"])
            assert torch.equal(z, z0) and torch.equal(p.predict_duration(dev["text_emb"], z), d0)
        p.set_option("use_pdl", 0)
        for a, b in zip(ref, _battery(p)):
            assert torch.equal(a, b)
        p.set_option("use_graph", 0)
        z = p.sample_style(dev["text_emb"], dev["prompt_feats"], 4, 2.0, noise=dev["noise"])
        assert torch.equal(z, z0)
        p.set_option("use_pdl", 1)
        z = p.sample_style(dev["text_emb"], dev["prompt_feats"], 4, 2.0, noise=dev["noise"])
        assert torch.equal(z, z0)
        # one TMA box per attention operand (permuted tensor map) vs one box per CFG branch: the same bytes land in the same places
        p.set_option("use_graph", 1)
        p.set_option("attn_box2", 0)
        for a, b in zip(ref, _battery(p)):
            assert torch.equal(a, b)
        p.set_option("attn_box2", 1)
    finally:
        p.close()
