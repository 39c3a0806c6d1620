"""One configuration of the path, a few calls (profiling driver for ncu launch lists): python tools/run_cfg.py B T steps sampler [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import styletts_zs_b200 as stz

B, T, steps, sampler = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
cfg = stz.DEFAULT
p = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
kind = {"teacher": stz.SAMPLER_TEACHER, "guided": stz.SAMPLER_GUIDED}.get(sampler, stz.SAMPLER_STUDENT)
inp = stz.synthetic_inputs(cfg, B, T, steps=steps, sampler=kind, seed=1234)
dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}
for i in range(reps):
    if i == reps - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    z = p.sample_style(dev["text_emb"], dev["prompt_feats"], steps, 2.0, noise=dev["noise"], sampler=sampler)
    if i == reps - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
print("ok", float(z.std()))
