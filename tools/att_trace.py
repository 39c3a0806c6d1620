"""Phase timeline of attention_tc_kernel (debug): clock64 deltas per CTA for the LAST attention launch of a sample_style call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import styletts_zs_b200 as stz

cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
B, T, steps = 64, 64, 1
inp = stz.synthetic_inputs(cfg, B, T, steps=steps, seed=1234)
dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}
tr = torch.zeros(296 * 32, dtype=torch.int64, device="cuda")
names = ["start", "alloc", "pdl", "produce", "landed+sync", "S done", "pass1+sync", "pass2", "fence+sync", "PV issued", "O done", "epi+sync",
         "produce2", "landed2", "S2", "pass1", "pass2", "sync", "PVi", "O2", "epi2"]
for k, v in [kv.split("=") for kv in os.environ.get("OPTS", "").split(",") if kv]:
    path.set_option(k, int(v))
    print("option", k, v)
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
    for cta in (0, 1, 100, 200, 250, 295):
        row = t[cta]
        d = [int(row[i] - row[0]) for i in range(21)]
        print(f"  cta {cta:3d}: " + "  ".join(f"{n}={v}" for n, v in zip(names, d) if v > 0))
    ex = (t[:216, 24:30] - t[:216, 2:3]).float().mean(0)
    print("  produce detail (cycles after the pdl stamp): unit0 [enter, tma issued, vis stored], next unit [enter, tma issued, vis stored]:", [int(x) for x in ex])
    two = t[:216]
    dd = (two[:, 1:21] - two[:, 0:20]).float().mean(0)
    print("  mean phase cycles (CTAs with 2 units):", {names[i + 1]: int(dd[i]) for i in range(20)})
    print("  mean total:", float((two[:, 20] - two[:, 0]).float().mean()))
