"""Phase timeline of attention_tc4_kernel (trace build): clock64 deltas per CTA for the LAST attention launch of a sample_style call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import styletts_zs_b200 as stz

cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
B, T, steps = int(os.environ.get("B", 64)), 64, 1
inp = stz.synthetic_inputs(cfg, B, T, steps=steps, seed=1234)
dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}
tr = torch.zeros(296 * 32, dtype=torch.int64, device="cuda")
names = ["start", "alloc", "pdl", "produced", "landed+sync", "S done", "pass1", "pass2", "P fenced+PV issued", "O done", "epi+sync"]
path.set_option("attn_ctas", 4)
for which, bits in (("cross-attention", 0), ("self-attention", 2)):
    path.set_option("ablate", bits)
    for _ in range(3):
        path.sample_style(dev["text_emb"], dev["prompt_feats"], steps, 2.0, noise=dev["noise"])
    torch.cuda.synchronize()
    path.lib.stz_debug_set_att_trace(path._h, C.c_void_p(tr.data_ptr()))
    path.sample_style(dev["text_emb"], dev["prompt_feats"], steps, 2.0, noise=dev["noise"])
    torch.cuda.synchronize()
    path.lib.stz_debug_set_att_trace(path._h, None)
    t = tr.view(296, 32).cpu()
    print(which)
    for cta in (0, 1, 100, 200, 295):
        row = t[cta]
        print(f"  cta {cta:3d}: " + "  ".join(f"{n}={int(row[i] - row[0])}" for i, n in enumerate(names)))
    dd = (t[:, 1:11] - t[:, 0:10]).float().mean(0)
    print("  mean phase cycles:", {names[i + 1]: int(dd[i]) for i in range(10)})
    print("  mean total:", float((t[:, 10] - t[:, 0]).float().mean()), " max end - min start:", int(t[:, 10].max() - t[:, 0].min()))
