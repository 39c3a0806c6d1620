"""Per-step timeline of lstm_tc_kernel (debug; build with `make -C styletts-zs_b200/csrc TRACE=1 -B`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import styletts_zs_b200 as stz

cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
B, T = 64, 64
inp = stz.synthetic_inputs(cfg, B, T, steps=1, seed=7)
style = (0.7 * torch.randn(B, cfg.n_style, cfg.d_style, generator=torch.Generator().manual_seed(5))).cuda()
te = inp["text_emb"].cuda()
tr = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
for _ in range(2):
    path.predict_duration(te, style)
torch.cuda.synchronize()
path.lib.stz_debug_set_lstm_trace(path._h, C.c_void_p(tr.data_ptr()))
path.predict_duration(te, style)
torch.cuda.synchronize()
path.lib.stz_debug_set_lstm_trace(path._h, None)
t = tr.view(64, 8).cpu()
names = ["h landed", "MMAs issued", "acc ready", "pre_s + sync", "st.async done"]
for s in (1, 2, 10, 30, 60):
    r = t[s]
    print(f"step {s}: " + "  ".join(f"{n} +{int(r[i + 1] - r[i])}" for i, n in enumerate(names)) + f"   step period {int(t[s + 1][0] - r[0]) if s < 63 else -1}")
d = (t[2:62, 1:6] - t[2:62, 0:5]).float().mean(0)
print("mean:", {n: int(d[i]) for i, n in enumerate(names)}, "period", float((t[3:62, 0] - t[2:61, 0]).float().mean()))
