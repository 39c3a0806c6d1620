"""fuse_ln variants (0 = GEMM + ln_mod kernels, 1 = one-CTA fused, 2 = cluster-of-two fused): agreement and time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import styletts_zs_b200 as stz

cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
B, T, steps = int(os.environ.get("B", 64)), 64, 4
inp = stz.synthetic_inputs(cfg, B, T, steps=steps, seed=1234)
dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}
samp = lambda: path.sample_style(dev["text_emb"], dev["prompt_feats"], steps, 2.0, noise=dev["noise"])


def timeit(fn, n=20):
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


ref = None
for mode in [int(x) for x in os.environ.get("MODES", "0,3").split(",")]:
    path.set_option("fuse_ln", mode)
    z = samp()
    torch.cuda.synchronize()
    if ref is None:
        ref = z
    err = float((z - ref).abs().max() / ref.abs().max())
    print(f"fuse_ln {mode}: finite {bool(torch.isfinite(z).all())} rel diff vs first {err:.3e}  sample_style {timeit(samp):.3f} ms", flush=True)
