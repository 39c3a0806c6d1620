"""GPU bring-up diagnostics: runs each stage in its own subprocess (a faulting kernel must not take
the later stages down) and prints per-stage parity numbers.  Usage on the GPU box:
    python tools/gpu_diag.py [stage ...]      (stages: gemm taps sample predictor timing)
Checker = oracle/ (test infrastructure); nothing here is product code.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-12))


def stage_gemm():
    import torch
    from styletts_zs_b200.path import op_gemm_bf16
    torch.manual_seed(0)
    out = {}
    for (M, N, K) in [(128, 128, 64), (128, 128, 256), (256, 256, 512), (100, 512, 512), (6400, 512, 512),
                      (6400, 2048, 512), (6400, 512, 2048), (2, 37888, 512), (1, 128, 64), (4096, 1024, 512)]:
        A = torch.randn(M, K, device="cuda").bfloat16()
        W = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
        b = torch.randn(N, device="cuda")
        ref = A.float() @ W.float().t() + b
        for impl in (1, 0):
            try:
                c = op_gemm_bf16(A, W, b, impl)
                torch.cuda.synchronize()
                out[f"{M}x{N}x{K}/impl{impl}"] = rel(c, ref)
            except Exception as e:  # noqa
                out[f"{M}x{N}x{K}/impl{impl}"] = f"ERR {e}"
        print(M, N, K, {k: v for k, v in out.items() if k.startswith(f"{M}x{N}x{K}/")}, flush=True)
    return out


def _setup(B=2, T=64, steps=1, sampler=0, seed=1234, var_len=None):
    import torch
    import styletts_zs_b200 as stz
    from oracle.model import OraclePath
    cfg = stz.DEFAULT
    w = stz.init_weights(cfg, 0)
    inp = stz.synthetic_inputs(cfg, B, T, steps=steps, sampler=sampler, seed=seed, var_len=var_len)
    return cfg, w, inp, stz, OraclePath


def stage_taps():
    """Residual stream after each sub-layer of eval 0 vs the bf16-emulating oracle."""
    import torch
    import oracle.model as om
    cfg, w, inp, stz, OraclePath = _setup()
    B, K, d = 2, cfg.n_style, cfg.d_model
    # record oracle taps by wrapping ops.lin on the residual adds: simplest is to re-run truncated nets
    path = stz.StyleTTSZSPath(cfg, w)
    out = {}
    for impl in (1, 0):
        path.set_option("gemm_impl", impl)
        path.set_option("use_graph", 0)
        for layer, stage in [(0, 0), (0, 1), (0, 2), (3, 2), (7, 2), (8, 0)]:
            n = B * K * (cfg.d_style if layer == cfg.n_layers else 2 * d)
            buf = torch.zeros(n, device="cuda")
            path.set_tap(0, layer, stage, buf)
            path.sample_style(inp["text_emb"], inp["prompt_feats"], 1, 2.0, noise=inp["noise"])
            torch.cuda.synchronize()
            path.set_tap(-1, -1, -1, None)
            got = buf.cpu()
            ref = oracle_tap(cfg, w, inp, layer, stage, emulate=True)
            ref32 = oracle_tap(cfg, w, inp, layer, stage, emulate=False)
            if layer == cfg.n_layers:
                got = got.view(B, K, cfg.d_style)
            else:
                got = got.view(B, K, 2, d).permute(2, 0, 1, 3)  # [branch, B, K, d]
            out[f"impl{impl}/L{layer}s{stage}"] = (rel(got, ref), rel(got, ref32), bool(torch.isfinite(got).all()))
            print(f"impl{impl} tap L{layer} s{stage}: vs emu {rel(got, ref):.3e}  vs fp32 {rel(got, ref32):.3e}", flush=True)
    return out


def oracle_tap(cfg, w, inp, layer, stage, emulate):
    """h after (layer, stage) of eval 0 for both branches [2,B,K,d]; layer == L -> guided F [B,K,Ds]."""
    import torch
    import oracle.model as om
    from oracle import schedule as S
    from styletts_zs_b200.spec import view_weights
    W = view_weights(cfg, w)
    ops = om._Ops(emulate)
    B, T, _ = inp["text_emb"].shape
    cond = om.Conditioning(cfg, W, inp["text_emb"], inp["text_mask"], inp["prompt_feats"], inp["prompt_mask"], ops)
    sig = S.student_sigmas(1, cfg)[0]
    c_skip, c_out, c_in, c_noise = S.edm_precond(sig, cfg.sigma_data)
    x = sig * inp["noise"][0]
    xin = c_in * x
    if layer == cfg.n_layers:
        Fc = om.denoiser_F(cfg, W, xin, c_noise, cond, 0, ops)
        Fu = om.denoiser_F(cfg, W, xin, c_noise, cond, 1, ops)
        return Fu + 2.0 * (Fc - Fu)
    outs = []
    for br in (0, 1):
        outs.append(_partial_F(cfg, W, xin, c_noise, cond, br, ops, layer, stage))
    return torch.stack(outs)


def _partial_F(cfg, W, x_in, c_noise, cond, branch, ops, stop_layer, stop_stage):
    import torch
    import torch.nn.functional as F
    import oracle.model as om
    from oracle import schedule as S
    d, L, H = cfg.d_model, cfg.n_layers, cfg.n_heads
    feat = torch.tensor(S.time_features(c_noise, cfg.d_time), dtype=torch.float64).to(torch.float32)
    t = F.linear(F.silu(F.linear(feat, W["time.w1"], W["time.b1"])), W["time.w2"], W["time.b2"])
    c = F.silu(t[None, :] + cond.pooled[branch])
    mod = ops.lin(c, W["mod.w"], W["mod.b"])
    m = lambda i: mod[:, i * d:(i + 1) * d][:, None, :]
    h = ops.lin(x_in, W["in.w"], W["in.b"]) + W["pos"][None]
    for l in range(L):
        p, o = f"l{l}.", 9 * l
        u = om.layer_norm(h) * (1 + m(o + 1)) + m(o + 0)
        qkv = ops.lin(u, W[p + "qkv.w"], W[p + "qkv.b"])
        a = om.attention(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], H, None, ops)
        h = h + m(o + 2) * ops.lin(a, W[p + "o.w"], W[p + "o.b"])
        if (l, 0) == (stop_layer, stop_stage):
            return h
        u = om.layer_norm(h) * (1 + m(o + 4)) + m(o + 3)
        q = ops.lin(u, W[p + "q2.w"], W[p + "q2.b"])
        kv = cond.kv[l][branch]
        a = om.attention(q, kv[..., :d], kv[..., d:], H, cond.ctx_mask[branch], ops)
        h = h + m(o + 5) * ops.lin(a, W[p + "o2.w"], W[p + "o2.b"])
        if (l, 1) == (stop_layer, stop_stage):
            return h
        u = om.layer_norm(h) * (1 + m(o + 7)) + m(o + 6)
        f = om.gelu_tanh(ops.lin(u, W[p + "ff1.w"], W[p + "ff1.b"]))
        h = h + m(o + 8) * ops.lin(f, W[p + "ff2.w"], W[p + "ff2.b"])
        if (l, 2) == (stop_layer, stop_stage):
            return h
    return h


def stage_sample():
    import torch
    out = {}
    for name, kw, B, T, steps, sampler in [
        ("student1", {}, 2, 64, 1, 0), ("student4", {}, 2, 64, 4, 0), ("teacher3", {}, 2, 64, 3, 1),
        ("student4_varlen", {"var_len": (16, 96)}, 5, 96, 4, 0), ("student1_B1", {}, 1, 64, 1, 0),
    ]:
        cfg, w, inp, stz, OraclePath = _setup(B, T, steps, sampler, **kw)
        ref = OraclePath(cfg, w).sample_style(inp["text_emb"], inp["prompt_feats"], steps, 2.0, text_mask=inp["text_mask"],
                                              noise=inp["noise"], sampler=sampler)
        path = stz.StyleTTSZSPath(cfg, w)
        for impl, graph in [(1, 0), (0, 0), (0, 1), (0, 1)]:
            path.set_option("gemm_impl", impl)
            path.set_option("use_graph", graph)
            try:
                z = path.sample_style(inp["text_emb"], inp["prompt_feats"], steps, 2.0, text_mask=inp["text_mask"],
                                      noise=inp["noise"], sampler=sampler)
                torch.cuda.synchronize()
                r = rel(z.cpu(), ref)
            except Exception as e:  # noqa
                r = f"ERR {e}"
            out[f"{name}/impl{impl}/graph{graph}"] = r
            print(name, "impl", impl, "graph", graph, "rel err vs fp32 oracle:", r, flush=True)
        path.close()
    return out


def stage_predictor():
    import torch
    out = {}
    for name, B, T, var in [("B2T64", 2, 64, None), ("B9T96var", 9, 96, (5, 96)), ("B1T16", 1, 16, None)]:
        cfg, w, inp, stz, OraclePath = _setup(B, T, 1, 0, var_len=var)
        o = OraclePath(cfg, w)
        style = 0.7 * torch.randn(B, cfg.n_style, cfg.d_style, generator=torch.Generator().manual_seed(5))
        dref, sref = o.predict_duration(inp["text_emb"], style, text_mask=inp["text_mask"], return_presum=True)
        path = stz.StyleTTSZSPath(cfg, w)
        d, s = path.predict_duration(inp["text_emb"], style, text_mask=inp["text_mask"], return_presum=True)
        torch.cuda.synchronize()
        d, s = d.cpu(), s.cpu()
        m = inp["text_mask"]
        agree = float((d[m] == dref[m]).float().mean())
        out[name] = dict(agree=agree, presum_maxabs=float((s[m] - sref[m]).abs().max()), pad_zero=bool((d[~m] == 0).all()),
                         dur_min=int(dref[m].min()), dur_max=int(dref[m].max()))
        print(name, out[name], flush=True)
        path.close()
    return out


def stage_timing():
    import torch
    cfg, w, inp, stz, _ = _setup(64, 64, 4, 0)
    path = stz.StyleTTSZSPath(cfg, w)
    dev = {k: v.cuda() for k, v in inp.items() if k != "lens"}
    out = {}
    for graph in (0, 1):
        path.set_option("use_graph", graph)
        for _ in range(3):
            z = path.sample_style(dev["text_emb"], dev["prompt_feats"], 4, 2.0, noise=dev["noise"])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            z = path.sample_style(dev["text_emb"], dev["prompt_feats"], 4, 2.0, noise=dev["noise"])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out[f"cfg2_graph{graph}_ms"] = ms
        print(f"cfg2 sample_style graph={graph}: {ms:.3f} ms  -> {64 / ms * 1e3:.0f} utt/s", flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d = path.predict_duration(dev["text_emb"], z)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        d = path.predict_duration(dev["text_emb"], z)
    e1.record()
    torch.cuda.synchronize()
    out["cfg2_predictor_ms"] = e0.elapsed_time(e1) / 5
    print("predictor B=64 T=64 ms:", out["cfg2_predictor_ms"], "launches", path.launch_count(), flush=True)
    return out


STAGES = dict(gemm=stage_gemm, taps=stage_taps, sample=stage_sample, predictor=stage_predictor, timing=stage_timing)

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        res = STAGES[sys.argv[2]]()
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"diag_{sys.argv[2]}.json"), "w") as f:
            json.dump(res, f, indent=1, default=str)
        sys.exit(0)
    stages = sys.argv[1:] or list(STAGES)
    for s in stages:
        t = time.time()
        print(f"===== stage {s}", flush=True)
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", s], timeout=600)
            print(f"===== stage {s} exit {p.returncode} in {time.time() - t:.1f}s", flush=True)
        except subprocess.TimeoutExpired:
            print(f"===== stage {s} TIMEOUT", flush=True)
