"""Is the GEMM mainloop bound per SM (shared memory / TMA engine / MMA) or by a shared resource (L2 -> SM bandwidth)?
One tile per CTA, 72 vs 144 busy CTAs, long K so that the mainloop dominates; per-tile time constant -> per-SM bound,
growing with the number of busy CTAs -> shared-resource bound."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import styletts_zs_b200 as stz
cfg = stz.DEFAULT
p = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
for bn in (256, 128):
    p.set_option("gemm_bn", bn)
    for K in (512, 2048, 8192):
        row = []
        for tiles_m in (2, 4, 9, 18, 36):
            M, N = 128 * tiles_m, 2048
            us = p.bench_gemm(M, N, K, 2, 30)
            ntiles = tiles_m * (N // bn)
            gb = ntiles * (128 + bn) * K * 2 / 1e9
            row.append(f"{ntiles:4d} tiles {us:7.2f} us {2.0 * M * N * K / us * 1e-6:7.1f} TF/s {gb / (us * 1e-6) / 1e3:5.2f} TB/s(L2->SM)")
        print(f"BN {bn} K {K}: " + " | ".join(row), flush=True)
