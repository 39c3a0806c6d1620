"""Reduced pass over the hot path for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_subset.py
Covers: the fused GEMM + AdaLN kernel at the benched row-tile count (cfg2 batch, 1 step), every GEMM epilogue, resident and
streaming tcgen05 attention (+ the mma.sync fallback), the cluster BiLSTM recurrence (DSMEM st.async), the pipelined host
slots, the guidance-conditioned student, the prosody heads, the regulator and the Philox generator.  No oracle here: the
parity tests are the correctness gate; this run only has to be hazard-free."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import styletts_zs_b200 as stz

cfg = stz.DEFAULT
p = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
small = os.environ.get("SMALL", "0") == "1"
ok = lambda t: bool(torch.isfinite(t.float()).all())

# cfg2 batch -> 50 row tiles -> gemmln3_kernel; 1 step keeps the sanitizer run short
B = 48 if small else 64
inp = stz.synthetic_inputs(cfg, B, 64, steps=1, seed=1)
z = p.sample_style(inp["text_emb"], inp["prompt_feats"], 1, 2.0, noise=inp["noise"])
assert p.get_option("last_fuse_mode") == 3 and ok(z)
d = p.predict_duration(inp["text_emb"], z)
print("cfg2-size fused path ok", flush=True)
# streaming attention + masks + teacher (per-eval path, ancestral noise)
inp = stz.synthetic_inputs(cfg, 2, 300, steps=2, sampler=stz.SAMPLER_TEACHER, seed=2, var_len=(100, 300))
z = p.sample_style(inp["text_emb"], inp["prompt_feats"], 2, 2.0, text_mask=inp["text_mask"], noise=inp["noise"], sampler="teacher")
assert ok(z)
d = p.predict_duration(inp["text_emb"], z, text_mask=inp["text_mask"])
print("streaming attention / teacher ok", flush=True)
# mma.sync fallback (long prompt)
inp = stz.synthetic_inputs(cfg, 2, 20, P=140, steps=1, seed=3)
assert ok(p.sample_style(inp["text_emb"], inp["prompt_feats"], 1, 2.0, noise=inp["noise"]))
# guided student, fused kernel forced
inp = stz.synthetic_inputs(cfg, 3, 40, steps=2, seed=4, var_len=(10, 40))
p.set_option("fuse_ln", 4)
assert ok(p.sample_style(inp["text_emb"], inp["prompt_feats"], 2, 2.0, text_mask=inp["text_mask"], noise=inp["noise"], sampler="guided"))
p.set_option("fuse_ln", 3)
print("guided student ok", flush=True)
# prosody heads + regulator
style = 0.7 * torch.randn(3, cfg.n_style, cfg.d_style)
f0, en, fl, dur = p.predict_prosody(inp["text_emb"], style, text_mask=inp["text_mask"], max_frames=400)
assert ok(f0) and ok(en)
fr, ln = p.regulate_length(torch.randn(3, 40, 64), dur, max_frames=300)
print("prosody heads ok", flush=True)
# pipelined host slots + device noise
a = stz.synthetic_inputs(cfg, 4, 32, steps=2, seed=5)
pin = {k: v.pin_memory() for k, v in a.items() if torch.is_tensor(v) and v.dtype == torch.float32}
o0 = p.synthesize_host(pin["text_emb"], pin["prompt_feats"], 2, 2.0, noise=pin["noise"], slot=0)
o1 = p.synthesize_host(pin["text_emb"], pin["prompt_feats"], 2, 2.0, seed=7, slot=1)
p.synthesize_host_wait(0); p.synthesize_host_wait(1)
assert ok(o0[0]) and ok(o1[0])
print("host pipeline ok", flush=True)
# epilogue unit entries
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(300, 512, device="cuda", generator=g).bfloat16()
W = (torch.randn(512, 512, device="cuda", generator=g) / 22).bfloat16()
bias = torch.randn(512, device="cuda", generator=g)
mod = torch.randn(8, 1536, device="cuda", generator=g)
h = torch.randn(300, 512, device="cuda", generator=g)
p.op_gemm_epi(A, W, bias, 3)
p.op_gemm_epi(A, W, bias, 4, out=h, mod=mod, gate_off=0)
p.op_gemm_ln(A, W, bias, h, mod, gate_off=0, shift_off=512, scale_off=1024)
torch.cuda.synchronize()
p.close()
print("sanitize subset done", flush=True)
