"""Timeline of gemmln3_kernel (debug; make TRACE=1): stamps of epilogue warp 2 of a few CTAs for the LAST fused launch of a call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import styletts_zs_b200 as stz
cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
path.set_option("fuse_ln", 4)  # force the fused kernel
path.set_option("use_graph", 0)
inp = stz.synthetic_inputs(cfg, 64, 64, steps=1, seed=1)
dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}
tr = torch.zeros(2 * 148 * 64, dtype=torch.int64, device="cuda")
for _ in range(2):
    path.sample_style(dev["text_emb"], dev["prompt_feats"], 1, 2.0, noise=dev["noise"])
torch.cuda.synchronize()
path.lib.stz_debug_set_gemm_trace(path._h, C.c_void_p(tr.data_ptr()))
path.sample_style(dev["text_emb"], dev["prompt_feats"], 1, 2.0, noise=dev["noise"])
torch.cuda.synchronize()
path.lib.stz_debug_set_gemm_trace(path._h, None)
t = tr.view(296, 64).cpu()[148:]
names = ["pdl", "acc ready", "h tile ready", "pass 1", "stats+bar", "cluster sync", "pass 2", "bar", "end"]
for cta in (0, 1, 50, 99):
    r = t[cta]
    print(f"   pass-1 chunk 0: tmem_ld {int(r[17]-r[16])} body {int(r[18]-r[17])}; chunk 1: tmem_ld {int(r[21]-r[20])} body {int(r[22]-r[21])}; between {int(r[20]-r[18])}; tr3->chunk0 {int(r[16]-r[3])}; chunk1 end->tr4 {int(r[4]-r[22])}")
    print(f"cta {cta}: " + "  ".join(f"{n} +{int(r[i + 1] - r[i])}" for i, n in enumerate(names)) + f"  total {int(r[9] - r[0])}")
