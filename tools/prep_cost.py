"""Fixed (per-call) vs per-evaluation cost of sample_style: linear fit over the step count."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import styletts_zs_b200 as stz

cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
B, T = int(os.environ.get("B", 64)), int(os.environ.get("T", 64))
inp = stz.synthetic_inputs(cfg, B, T, steps=8, seed=1234)
dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}


def t(steps, n=30):
    f = lambda: path.sample_style(dev["text_emb"], dev["prompt_feats"], steps, 2.0, noise=dev["noise"])
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


ts = {s: t(s) for s in (1, 2, 4, 8)}
per_eval = (ts[8] - ts[1]) / 7
print({k: round(v, 4) for k, v in ts.items()}, "per eval ms", round(per_eval, 4), "fixed ms", round(ts[1] - per_eval, 4))
path.set_option("profile", 1)
path.sample_style(dev["text_emb"], dev["prompt_feats"], 1, 2.0, noise=dev["noise"])
print(path.profile_read())
