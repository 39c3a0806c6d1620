"""On-device noise generator measurement (supplementary; not the bench line): achieved GB/s of philox_normal_kernel.
Algorithmic bytes per launch = 4 * elements written (a pure writer: 16 B per Philox counter)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import styletts_zs_b200 as stz

cfg = stz.DEFAULT
peak = 6540.5
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
n = cfg.n_style * cfg.d_style
res = []
for name, slices, B in [("cfg2 student (1 slice, B=64)", 1, 64), ("cfg3 teacher (33 slices, B=32)", 33, 32), ("cfg4 student (1 slice, B=256)", 1, 256),
                        ("large (33 slices, B=256)", 33, 256)]:
    out = torch.empty(slices, B, n, dtype=torch.float32, device="cuda")
    for _ in range(3):
        stz.philox_normal(1234, 0, slices, B, n, out=out)
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); stz.philox_normal(1234, 0, slices, B, n, out=out); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    nbytes = 4.0 * out.numel()
    res.append({"case": name, "elements": out.numel(), "ms": round(ms, 4), "algorithmic_bytes": nbytes,
                "achieved_GBps": round(nbytes / ms * 1e-6, 1), "peak_GBps": peak, "frac": round(nbytes / ms * 1e-6 / peak, 3),
                "gsamples_per_s": round(out.numel() / ms * 1e-6, 2)})
    print(res[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/philox.json", "w"), indent=1)
