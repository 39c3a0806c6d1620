import os, sys
sys.path.insert(0, "/root/repo")
import torch, styletts_zs_b200 as stz
cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
for bn in (0, 128, 256):
    path.set_option("gemm_bn", bn)
    for (M, N, K, epi) in [(7296, 8192, 512, 2), (4096, 2048, 1920, 2), (128, 37888, 512, 2), (512, 37888, 512, 2), (4096, 512, 512, 2), (3200, 512, 512, 2)]:
        us = path.bench_gemm(M, N, K, epi, 20)
        print("bn", bn, M, N, K, round(us, 1), "us", round(2.0 * M * N * K / us * 1e-6), "TF", flush=True)
