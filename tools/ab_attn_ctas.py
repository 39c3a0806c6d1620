"""A/B of the resident-key attention kernels: attention_tc4_kernel (4 CTAs/SM, one unit in flight) vs attention_tc2_kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import styletts_zs_b200 as stz
cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
for B, T, steps, sampler in ((64, 64, 4, "student"), (1, 64, 1, "student"), (32, 64, 8, "teacher"), (256, 64, 4, "student"), (64, 64, 4, "guided")):
    inp = stz.synthetic_inputs(cfg, B, T, steps=steps, seed=7, sampler={"student": stz.SAMPLER_STUDENT, "teacher": stz.SAMPLER_TEACHER, "guided": stz.SAMPLER_GUIDED}[sampler])
    dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}
    outs = {}
    for rnd in range(2):
        for ctas in (2, 4):
            path.set_option("attn_ctas", ctas)
            run = lambda: path.sample_style(dev["text_emb"], dev["prompt_feats"], steps, 2.0, noise=dev["noise"], sampler=sampler)
            for _ in range(3):
                z = run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(20):
                z = run()
            e1.record(); torch.cuda.synchronize()
            outs[ctas] = z.clone()
            if rnd:
                print(f"B {B:3d} T {T} {sampler:8s} x{steps}  attn_ctas {ctas}: sample_style {e0.elapsed_time(e1) / 20:.4f} ms")
    print("   max |diff| between the kernels:", float((outs[2].float() - outs[4].float()).abs().max()))
