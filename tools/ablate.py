"""In-graph cost attribution (not a bench): time cfg2's sample_style with one kernel family of the denoiser
evaluation removed from the captured graph at a time (results are wrong by construction; only the time matters).
The difference to the full graph is what that family costs INSIDE the graph, PDL overlap included."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import styletts_zs_b200 as stz

cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
B, T, steps = int(os.environ.get("B", 64)), int(os.environ.get("T", 64)), int(os.environ.get("STEPS", 4))
inp = stz.synthetic_inputs(cfg, B, T, steps=steps, seed=1234)
dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}


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


samp = lambda: path.sample_style(dev["text_emb"], dev["prompt_feats"], steps, 2.0, noise=dev["noise"])
z = samp()
pred = lambda: path.predict_duration(dev["text_emb"], z)
for ch in (1, 2, 3, 4, 6, 8):
    path.set_option("chains", ch)
    print(f"chains {ch}: sample_style {timeit(samp):.3f} ms", flush=True)
path.set_option("chains", int(os.environ.get("CHAINS", 1)))
if os.environ.get("SWEEP_ONLY"):
    sys.exit(0)
full = timeit(samp)
print(f"full sample_style {full:.3f} ms   predict_duration {timeit(pred):.3f} ms", flush=True)
names = {512: "attention bodies (empty kernels instead)", 1: "self-attn", 2: "cross-attn", 4: "ln_mod", 8: "qkv gemm", 16: "attn out-proj gemms (2/layer)", 32: "q2 gemm",
         64: "ff1 gemm", 128: "ff2 gemm", 256: "mod gemm", 3: "both attentions", 511 - 256: "everything per-layer",
         8 + 16 + 32 + 64 + 128: "all layer gemms"}
for bits, name in names.items():
    path.set_option("ablate", bits)
    t = timeit(samp)
    print(f"without {name:32s} {t:.3f} ms   (family costs {full - t:.3f} ms)", flush=True)
path.set_option("ablate", 0)
