"""lstm_impl variants (0 tcgen05, 2 fp32 FFMA cluster, 1 generic): duration agreement with the FFMA kernel / oracle-free timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import styletts_zs_b200 as stz

cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))


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


for (B, T, var) in [(64, 64, None), (5, 37, (9, 37)), (256, 512, (16, 512))]:
    inp = stz.synthetic_inputs(cfg, B, T, steps=1, seed=7, var_len=var)
    style = (0.7 * torch.randn(B, cfg.n_style, cfg.d_style, generator=torch.Generator().manual_seed(5))).cuda()
    te = inp["text_emb"].cuda()
    tm = inp["text_mask"].cuda() if var else None
    ref = None
    for impl in (2, 0):
        path.set_option("lstm_impl", impl)
        d, s = path.predict_duration(te, style, text_mask=tm, return_presum=True)
        torch.cuda.synchronize()
        if ref is None:
            ref = (d, s)
        m = inp["text_mask"].cuda() if var else torch.ones(B, T, dtype=torch.bool, device="cuda")
        agree = float((d[m] == ref[0][m]).float().mean())
        err = float((s[m] - ref[1][m]).abs().max())
        ms = timeit(lambda: path.predict_duration(te, style, text_mask=tm), n=10)
        print(f"B {B} T {T} lstm_impl {impl}: agree {agree:.5f} max |presum diff| {err:.3e} finite {bool(torch.isfinite(s).all())} predict_duration {ms:.3f} ms", flush=True)
