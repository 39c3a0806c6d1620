"""Phase timeline of gemm2_kernel (debug; build with `make -C styletts-zs_b200/csrc TRACE=1 -B`): clock64 stamps per CTA
of the LAST launch of a back-to-back series, plus a finer breakdown of the first epilogue chunks of warp 2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import styletts_zs_b200 as stz

cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
tr = torch.zeros(148 * 64, dtype=torch.int64, device="cuda")
M = 6400
path.set_option("gemm_cluster", int(os.environ.get("CLUSTER", 0)))
path.lib.stz_debug_set_gemm_dbg.argtypes = [C.c_void_p, C.c_int]
for dbg in [int(x) for x in os.environ.get("DBG", "0").split(",")]:
    path.lib.stz_debug_set_gemm_dbg(path._h, dbg)
    print("=== dbg flags", dbg)
    for name, N, K, epi in [("ffn1", 2048, 512, 3), ("qkv", 1536, 512, 2), ("attn_out", 512, 512, 4)]:
        tr.zero_()
        path.lib.stz_debug_set_gemm_trace(path._h, C.c_void_p(tr.data_ptr()))
        us = path.bench_gemm(M, N, K, epi, 20)
        torch.cuda.synchronize()
        path.lib.stz_debug_set_gemm_trace(path._h, None)
        t = tr.view(148, 64).cpu()
        print(f"{name}: M {M} N {N} K {K}  {us:.2f} us/launch (traced)")
        for cta in (0, 73, 99, 147):
            r = t[cta]
            if r[0] == 0:
                continue
            d = lambda i: int(r[i] - r[0]) if r[i] > 0 else -1
            tiles = "  ".join(f"[t{k}: land {d(8+4*k)} commit {d(9+4*k)} accrdy {d(10+4*k)} drained {d(11+4*k)}]"
                              for k in range(4) if r[8 + 4 * k] > 0)
            print(f"  cta {cta:3d}: alloc {d(1)} pdl {d(2)} tma0 {d(3)} tmaN {d(4)} end {d(5)}  {tiles}")
        r = t[73]
        for ci in range(2):
            e = [int(r[32 + 8 * ci + k] - r[0]) for k in range(7)]
            print(f"    cta 73 epilogue warp 2, chunk {ci}: start {e[0]}  tmem_ld+wait +{e[1]-e[0]}  math/prefetch +{e[2]-e[1]}  "
                  f"wait_read +{e[3]-e[2]}  st.shared +{e[4]-e[3]}  fence+syncwarp +{e[5]-e[4]}  tma issue +{e[6]-e[5]}")
