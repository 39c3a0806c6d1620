"""Supplementary measurements (not the bench line): the other BASELINE.json configs and the per-GPU batch sweep.
    python tools/sweep.py > gpurun_out/sweep.json
cfg1: B=1,T=64, student 1 step + predictor.  cfg3: B=32,T=64, teacher 32 ADPM2 steps (64 evals).  cfg4: B=256,
T in [16,512] (padded to 512, masks), student 4 steps + predictor.  cfg5: student 4 steps + predictor, T=64, B = 1..1024.
Device-resident inputs, CUDA events, best-of-N wall per call pair; utterances/s and path-RTF (80 frames/s)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import styletts_zs_b200 as stz

cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))


def run(B, T, steps, sampler, var_len=None, iters=5):
    kind = 1 if sampler == "teacher" else 0
    inp = stz.synthetic_inputs(cfg, B, T, steps=steps, sampler=kind, seed=1234, var_len=var_len)
    dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}
    tm = inp["text_mask"].cuda() if var_len else None

    def step():
        z = path.sample_style(dev["text_emb"], dev["prompt_feats"], steps, 2.0, noise=dev["noise"], text_mask=tm, sampler=sampler)
        return z, path.predict_duration(dev["text_emb"], z, text_mask=tm)
    for _ in range(2):
        z, d = step()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); z, d = step(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    frames = int(d.sum())
    return {"B": B, "T": T, "steps": steps, "sampler": sampler, "ms": round(best, 3), "utt_per_s": round(B / best * 1e3, 1),
            "path_rtf": best * 1e-3 / (frames / 80.0) if frames else None, "finite": bool(torch.isfinite(z).all())}


out = {"cfg1": run(1, 64, 1, "student"), "cfg2": run(64, 64, 4, "student"), "cfg3": run(32, 64, 32, "teacher", iters=3),
       "cfg4": run(256, 512, 4, "student", var_len=(16, 512), iters=3),
       "cfg5_batch_sweep": [run(B, 64, 4, "student", iters=3) for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024)]}
print(json.dumps(out, indent=1))
