import sys; sys.path.insert(0, "/root/repo")
import torch, styletts_zs_b200 as stz
cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
path.set_option("use_graph", 0)
inp = stz.synthetic_inputs(cfg, 1, 16, steps=1)
z = path.sample_style(inp["text_emb"], inp["prompt_feats"], 1, 2.0, noise=inp["noise"])
torch.cuda.synchronize()
print("ok", float(z.abs().max()))
