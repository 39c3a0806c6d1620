"""GPU tuning sweeps (not a bench): time sample_style / predict_duration of cfg2 under library knobs."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import styletts_zs_b200 as stz

cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
B, T, steps = int(os.environ.get("B", 64)), int(os.environ.get("T", 64)), 4
inp = stz.synthetic_inputs(cfg, B, T, steps=steps, seed=1234)
dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}


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


z = path.sample_style(dev["text_emb"], dev["prompt_feats"], steps, 2.0, noise=dev["noise"])
samp = lambda: path.sample_style(dev["text_emb"], dev["prompt_feats"], steps, 2.0, noise=dev["noise"])
pred = lambda: path.predict_duration(dev["text_emb"], z)
res = {}
for knobs in [{}, {"gemm_cluster": 1}, {"use_pdl": 0}]:
    for k, v in knobs.items():
        path.set_option(k, v)
    res[json.dumps(knobs)] = (round(timeit(samp), 3), round(timeit(pred), 3))
    print(knobs, "sample_style ms", res[json.dumps(knobs)][0], "predict_duration ms", res[json.dumps(knobs)][1], flush=True)
    for k in knobs:
        path.set_option(k, {"gemm_bn": 0, "use_pdl": 1, "use_graph": 1, "fuse_ln": 0, "gemm_cluster": 0}[k])
for name, N, K, epi in [("qkv", 1536, 512, 2), ("attn_out", 512, 512, 4), ("q_cross", 512, 512, 2), ("ffn1", 2048, 512, 3), ("ffn2", 512, 2048, 4)]:
    R = 2 * B * cfg.n_style
    out = {}
    for cl in (1, 0):
        path.set_option("gemm_cluster", cl)
        us = path.bench_gemm(R, N, K, epi, 50)
        out[cl] = (round(us, 2), round(2.0 * R * N * K / us * 1e-6, 1))
    path.set_option("gemm_cluster", 0)
    print(name, "M", R, "N", N, "K", K, "cluster:", out[1], "no cluster:", out[0], "(us, TFLOP/s)", flush=True)
for bn in (128, 256):
    path.set_option("gemm_bn", bn)
    for name, N, K, epi in [("attn_out", 512, 512, 4), ("q_cross", 512, 512, 2), ("ffn2", 512, 2048, 4)]:
        R = 2 * B * cfg.n_style
        us = path.bench_gemm(R, N, K, epi, 50)
        print("gemm_bn", bn, name, round(us, 2), "us", flush=True)
    print("gemm_bn", bn, "sample_style ms", round(timeit(samp), 3), flush=True)
path.set_option("gemm_bn", 0)
