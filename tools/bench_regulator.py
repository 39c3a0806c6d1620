"""Length regulator measurement (supplementary; not the bench line): achieved HBM GB/s at cfg2 / cfg4 sizes.
Algorithmic bytes per launch = 4*C*(frames written: B*F_max) + 4*C*(distinct feature rows read: B*T) + 4*B*T (durations)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import styletts_zs_b200 as stz

cfg = stz.DEFAULT
path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0))
peak = 6540.5
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = []
for (B, T, C, F) in [(64, 64, 640, 1024), (256, 512, 640, 4096), (1024, 64, 640, 1024)]:
    g = torch.Generator().manual_seed(1)
    feats = torch.randn(B, T, C, generator=g).cuda()
    dur = torch.randint(1, 16, (B, T), generator=g, dtype=torch.int32).cuda()
    for _ in range(3):
        path.regulate_length(feats, dur, max_frames=F)
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fr, ln = path.regulate_length(feats, dur, max_frames=F); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    nbytes = 4.0 * C * B * F + 4.0 * C * B * T + 4.0 * B * T
    out.append({"B": B, "T": T, "C": C, "F_max": F, "ms": round(ms, 4), "frames": int(ln.sum()), "algorithmic_bytes": nbytes,
                "achieved_GBps": round(nbytes / ms * 1e-6, 1), "peak_GBps": peak, "frac": round(nbytes / ms * 1e-6 / peak, 3),
                "note": "includes the torch.empty of the outputs and the ctypes call; L2 flushed before each timed call"})
    print(out[-1], flush=True)
json.dump(out, open("gpurun_out/regulator.json", "w"), indent=1)
