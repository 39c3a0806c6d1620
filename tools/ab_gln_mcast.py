"""A/B (needs tools/experiments/gemmln3_a_multicast.diff.txt applied): A tiles of the fused kernel's CTA pair TMA-multicast to both CTAs (knob gln_mcast)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import styletts_zs_b200 as stz
cfg = stz.DEFAULT
p = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))

# correctness first: the kernel-level entry against fp32 torch, multicast on
g = torch.Generator(device="cuda").manual_seed(3)
for M, K in ((6400, 512), (6400, 2048), (1000, 512)):
    N = 512
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    nseq = (M + 49) // 50 + 2
    mod = 0.5 * torch.randn(nseq, 4 * N, device="cuda", generator=g)
    h0 = torch.randn(M, N, device="cuda", generator=g)
    outs = {}
    for mc in (0, 1):
        p.set_option("gln_mcast", mc)
        h = h0.clone()
        u = p.op_gemm_ln(A, W, b, h, mod, mode=0, gate_off=0, shift_off=N, scale_off=2 * N)
        torch.cuda.synchronize()
        outs[mc] = (h, u)
    print(f"M {M} K {K}: multicast == plain: h {torch.equal(outs[0][0], outs[1][0])} u {torch.equal(outs[0][1], outs[1][1])}", flush=True)


def timeit(fn, n=40):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for B in (64, 48, 72):
    inp = stz.synthetic_inputs(cfg, B, 64, steps=4, seed=1234)
    dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}
    samp = lambda: p.sample_style(dev["text_emb"], dev["prompt_feats"], 4, 2.0, noise=dev["noise"])
    res = {0: [], 1: []}
    zs = {}
    for rnd in range(3):
        for mc in (0, 1):
            p.set_option("gln_mcast", mc)
            res[mc].append(round(timeit(samp), 4))
            zs[mc] = samp().clone()
    print(f"B {B}: sample_style ms plain {res[0]}  multicast {res[1]}  max|dz| {float((zs[0] - zs[1]).abs().max()):.2e}", flush=True)
