"""Attribution: cost of the cross-attention K/V HBM misses (ablate bit 1024: every layer reads layer 0's K/V)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import styletts_zs_b200 as stz

cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
B, T = int(os.environ.get("B", 64)), int(os.environ.get("T", 64))
inp = stz.synthetic_inputs(cfg, B, T, steps=4, seed=1234)
dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}


def t(n=30):
    f = lambda: path.sample_style(dev["text_emb"], dev["prompt_feats"], 4, 2.0, noise=dev["noise"])
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for ab in (0, 1024, 0, 1024, 2, 2 | 1024):
    path.set_option("ablate", ab)
    print("ablate", ab, "sample_style ms", round(t(), 4), flush=True)
