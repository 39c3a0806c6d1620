"""A/B of the cta_group::2 CTA-pair GEMM (experiments build: STZ_LIBRARY=.../libstz_exp.so): isolated shapes and cfg2 / B = 256 sample_style."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import styletts_zs_b200 as stz
cfg = stz.DEFAULT
p = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for cl in (0, 1):
    p.set_option("gemm_cluster", cl)
    row = []
    for name, M, N, K, epi in [("qkv", 6400, 1536, 512, 2), ("ffn1", 6400, 2048, 512, 3), ("kv_prep", 7296, 8192, 512, 2), ("ffn1 B=256", 25600, 2048, 512, 3),
                               ("qkv B=256", 25600, 1536, 512, 2), ("big", 8192, 8192, 8192, 2)]:
        us = p.bench_gemm(M, N, K, epi, 20)
        row.append(f"{name} {us:.2f} us {2.0 * M * N * K / us * 1e-6:.0f} TF/s")
    print(f"gemm_cluster {cl}: " + " | ".join(row), flush=True)
for B in (64, 256):
    inp = stz.synthetic_inputs(cfg, B, 64, steps=4, seed=1234)
    dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}
    for cl in (0, 1):
        p.set_option("gemm_cluster", cl)
        t = timeit(lambda: p.sample_style(dev["text_emb"], dev["prompt_feats"], 4, 2.0, noise=dev["noise"]))
        print(f"B {B} gemm_cluster {cl}: sample_style {t:.3f} ms", flush=True)
