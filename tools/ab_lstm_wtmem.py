"""A/B of the BiLSTM recurrence forms: W_hh in tensor memory (product) vs shared memory, 8 vs 16 sequences per cluster."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import styletts_zs_b200 as stz
cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
for B, T in ((64, 64), (128, 64), (256, 128), (512, 128)):
    inp = stz.synthetic_inputs(cfg, B, T, steps=1, seed=7)
    style = (0.7 * torch.randn(B, cfg.n_style, cfg.d_style, generator=torch.Generator().manual_seed(5))).cuda()
    te = inp["text_emb"].cuda()
    outs = {}
    forms = {"w_tmem nb16": {"lstm_nb": 16}, "w_tmem nb24": {"lstm_nb": 24}, "w_tmem auto": {}, "w_smem nb16": {"lstm_impl": 3}}
    for rnd in range(2):
        for name, kn in forms.items():
            path.set_option("lstm_nb", 0); path.set_option("lstm_impl", 0)
            for k, v in kn.items():
                path.set_option(k, v)
            for _ in range(3):
                d = path.predict_duration(te, style)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(20):
                d = path.predict_duration(te, style)
            e1.record(); torch.cuda.synchronize()
            outs[name] = d.clone()
            if rnd:
                print(f"B {B} T {T}  {name:12s} predict_duration {e0.elapsed_time(e1) / 20:.4f} ms")
    print("  identical durations:", all(torch.equal(outs["w_tmem nb16"], o) for o in outs.values()))
