"""A/B: rows per CTA pair of the fused GEMM + AdaLN kernel = smallest multiple of 8 that still fits one wave of pairs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import styletts_zs_b200 as stz
cfg = stz.DEFAULT
p = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))


def timeit(fn, n=10):
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


for B, sampler, steps in [(8, "student", 4), (16, "student", 4), (32, "student", 4), (32, "teacher", 8), (64, "guided", 4), (48, "student", 4), (56, "student", 4), (64, "student", 4), (72, "student", 4), (80, "student", 4), (94, "student", 4)]:
    kind = {"teacher": stz.SAMPLER_TEACHER, "guided": stz.SAMPLER_GUIDED}.get(sampler, stz.SAMPLER_STUDENT)
    inp = stz.synthetic_inputs(cfg, B, 64, steps=steps, sampler=kind, seed=1234)
    dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}
    samp = lambda: p.sample_style(dev["text_emb"], dev["prompt_feats"], steps, 2.0, noise=dev["noise"], sampler=sampler)
    nbr = 1 if sampler == "guided" else 2
    rows = nbr * B * cfg.n_style
    fit = max(8, -(-rows // 74 + 7) // 8 * 8 if False else ((rows + 73) // 74 + 7) // 8 * 8)
    res = {}
    for name, (fl, tr) in {"unfused": (0, 0), "auto": (3, 0), f"fused/{fit}": (4, fit), f"fused/{fit + 8}": (4, fit + 8), "fused/96": (4, 96), "fused/128": (4, 128)}.items():
        if tr > 128:
            continue
        p.set_option("fuse_ln", fl); p.set_option("gln_tile_rows", tr)
        res[name] = timeit(samp)
    print(f"B {B:4d} {sampler:8s} rows {rows:6d}: " + "  ".join(f"{k} {v:.3f}" for k, v in res.items()), flush=True)
