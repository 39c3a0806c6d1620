"""CPU experiment: which bf16-operand GEMMs dominate the error vs the fp32 oracle?  Runs the
bf16-emulating oracle with selected ops kept exact.  Diagnostic only."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import styletts_zs_b200 as stz
from oracle.model import OraclePath
torch.set_num_threads(8)
cfg = stz.DEFAULT
w = stz.init_weights(cfg, 0)
cases = {"teacher2": (2, 32, 2, 1, 1236, 2.0), "student1": (2, 32, 1, 0, 1234, 2.0), "student4": (2, 32, 4, 0, 1235, 2.0)}
sets = [(), ("in",), ("out",), ("in", "out"), ("in", "out", "mod"), ("mod",), ("in", "out", "mod", "kv2", "ctx"),
        ("ff1", "ff2"), ("qkv", "o", "q2", "o2"), ("in", "out", "mod", "ff2", "o", "o2")]
for name, (B, T, steps, sampler, seed, scale) in cases.items():
    inp = stz.synthetic_inputs(cfg, B, T, steps=steps, sampler=sampler, seed=seed)
    ref = OraclePath(cfg, w).sample_style(inp["text_emb"], inp["prompt_feats"], steps, scale, noise=inp["noise"], sampler=sampler)
    for ex in sets:
        z = OraclePath(cfg, w, True, ex).sample_style(inp["text_emb"], inp["prompt_feats"], steps, scale, noise=inp["noise"], sampler=sampler)
        e = float((z - ref).abs().max() / ref.abs().max())
        rms = float((z - ref).pow(2).mean().sqrt() / ref.abs().max())
        print(f"{name:10s} exact={','.join(ex) or '-':32s} max-rel {e:.3e}  rms/max {rms:.3e}", flush=True)
