"""What paces the GEMM mainloop?  (trace build: STZ_LIBRARY=.../libstz_trace.so)  One tile per CTA, 16 CTAs, K = 2048 vs 8192 ->
cycles per 64-wide k-block with: the full loop; no W loads; no A loads; no loads at all; one MMA instead of four."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import styletts_zs_b200 as stz
cfg = stz.DEFAULT
p = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
p.lib.stz_debug_set_gemm_dbg.argtypes = [C.c_void_p, C.c_int]
names = {0: "full loop", 8: "no W loads", 16: "no A loads", 24: "no loads", 32: "1 of 4 MMAs", 56: "no loads, 1 MMA"}
for bn in (256, 128):
    p.set_option("gemm_bn", bn)
    for tiles_m in (2, 18):
        for dbg, name in names.items():
            p.lib.stz_debug_set_gemm_dbg(p._h, dbg)
            t = {K: p.bench_gemm(128 * tiles_m, 2048, K, 2, 20) for K in (2048, 8192)}
            per_kb = (t[8192] - t[2048]) / 96 * 1e-6 * 1.965e9
            print(f"BN {bn} {tiles_m * 2048 // bn:4d} tiles  {name:16s}: K=2048 {t[2048]:7.2f} us  K=8192 {t[8192]:7.2f} us  -> {per_kb:6.0f} cycles per k-block (at 1965 MHz)", flush=True)
p.lib.stz_debug_set_gemm_dbg(p._h, 0)

# issue-loop timeline of the MMA thread (k-blocks 8..11 of the first tile of CTA 0): wait(full) | fence | 4 MMAs | commit
tr = torch.zeros(148 * 64, dtype=torch.int64, device="cuda")
for bn in (256, 128):
    p.set_option("gemm_bn", bn)
    for dbg in (0, 24):
        p.lib.stz_debug_set_gemm_dbg(p._h, dbg)
        tr.zero_()
        p.lib.stz_debug_set_gemm_trace(p._h, C.c_void_p(tr.data_ptr()))
        p.bench_gemm(256, 2048, 2048, 2, 5)
        torch.cuda.synchronize()
        p.lib.stz_debug_set_gemm_trace(p._h, None)
        r = tr.view(148, 64).cpu()[0]
        for i in range(4):
            e = [int(r[40 + 5 * i + j]) for j in range(5)]
            nxt = int(r[40 + 5 * (i + 1)]) if i < 3 else None
            print(f"BN {bn} dbg {dbg} k-block {8 + i}: wait {e[1]-e[0]:5d}  fence {e[2]-e[1]:4d}  mma x4 {e[3]-e[2]:5d}  commit {e[4]-e[3]:4d}" + (f"  loop-back {nxt - e[4]:4d}" if nxt else ""), flush=True)
p.lib.stz_debug_set_gemm_dbg(p._h, 0)
