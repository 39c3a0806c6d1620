import os, sys, torch
sys.path.insert(0, "/root/repo")
import styletts_zs_b200 as stz
cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
B, T = 512, 128
inp = stz.synthetic_inputs(cfg, B, T, steps=1, seed=7)
style = (0.7 * torch.randn(B, cfg.n_style, cfg.d_style, generator=torch.Generator().manual_seed(5))).cuda()
te = inp["text_emb"].cuda()
outs = {}
for rnd in range(2):
  for impl in (0, 3):
    path.set_option("lstm_impl", impl)
    for _ in range(3): d = path.predict_duration(te, style)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20): d = path.predict_duration(te, style)
    e1.record(); torch.cuda.synchronize()
    outs[impl] = [x.clone() if torch.is_tensor(x) else x for x in (d if isinstance(d, (tuple, list)) else [d])]
    print("lstm_impl", impl, "predict_duration ms", e0.elapsed_time(e1) / 20)
for a, b in zip(outs[0], outs[3]):
    if torch.is_tensor(a): print("equal", torch.equal(a, b), (a.float() - b.float()).abs().max().item())
