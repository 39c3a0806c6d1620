"""ORACLE (test infrastructure, not product code): fp32 PyTorch CPU restatement of the
StyleTTS-ZS inference hot path — style denoiser, CFG sampler loop, duration predictor.

PARITY UNPINNED: the reference has no implementation to follow (/root/reference/README.md:15-16);
this follows SURVEY.md §8(a) rows a-3 .. a-11, which restate /root/reference/README.md:5
("a diffusion model ... to sample this time-varying style code", "classifier-free guidance",
"distilled") with the StyleTTS-lineage shapes.  Pinned by analytic KATs (tests/test_oracle_kat.py)
and frozen fixtures (tests/golden/).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import schedule as S

SAMPLER_STUDENT, SAMPLER_TEACHER, SAMPLER_GUIDED = 0, 1, 2


def _bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


class _Ops:
    """GEMM operand precision.  emulate_bf16=True rounds tensor-core operands the way the CUDA
    path does (bf16 operands, fp32 accumulate) — a diagnostic for tests, never the reference."""

    def __init__(self, emulate_bf16: bool, exact_tags=()):
        self.emu = emulate_bf16
        self.exact_tags = set(exact_tags)   # ops the CUDA path runs in split-bf16 (fp32-grade) precision

    def lin(self, x, w, b=None, tag=None):
        if self.emu and tag not in self.exact_tags:
            x, w = _bf16(x), _bf16(w)
        return F.linear(x, w, b)

    def r(self, x):
        return _bf16(x) if self.emu else x


def layer_norm(x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, x.shape[-1:], eps=1e-5)


def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    return F.gelu(x, approximate="tanh")


def masked_mean(x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    m = mask.to(x.dtype)[..., None]
    return (x * m).sum(1) / m.sum(1).clamp(min=1.0)


def attention(q, k, v, n_heads: int, key_mask: Optional[torch.Tensor], ops: _Ops):
    """q [N,Lq,d], k/v [N,Lk,d], key_mask [N,Lk] bool (True = attend).  a-5."""
    N, Lq, d = q.shape
    dh = d // n_heads
    qh = ops.r(q).view(N, Lq, n_heads, dh).transpose(1, 2)
    kh = ops.r(k).view(N, -1, n_heads, dh).transpose(1, 2)
    vh = ops.r(v).view(N, -1, n_heads, dh).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) / math.sqrt(dh)
    if key_mask is not None:
        s = s.masked_fill(~key_mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    if ops.emu:
        # CUDA path: P is rounded to bf16 for the PV MMA, the row sum stays fp32
        e = torch.exp(s - s.amax(-1, keepdim=True))
        o = (_bf16(e) @ vh) / e.sum(-1, keepdim=True)
    else:
        o = p @ vh
    return o.transpose(1, 2).reshape(N, Lq, d)


class Conditioning:
    """a-3: everything that does not depend on the sampler state x."""

    def __init__(self, cfg, W: Dict[str, torch.Tensor], text_emb, text_mask, prompt_feats, prompt_mask,
                 ops: _Ops):
        B = text_emb.shape[0]
        self.B = B
        ct = layer_norm(ops.lin(text_emb, W["ctx_text.w"], W["ctx_text.b"], tag="ctx") + W["type_emb"][0])
        cp = layer_norm(ops.lin(prompt_feats, W["ctx_prompt.w"], W["ctx_prompt.b"], tag="ctx") + W["type_emb"][1])
        cn = layer_norm(W["null_tok"] + W["type_emb"][1]).view(1, 1, -1).expand(B, 1, -1)
        # cond branch: [text ; prompt], uncond branch: [text ; null token]
        self.ctx = (torch.cat([ct, cp], 1), torch.cat([ct, cn], 1))
        self.ctx_mask = (torch.cat([text_mask, prompt_mask], 1),
                         torch.cat([text_mask, torch.ones(B, 1, dtype=torch.bool)], 1))
        pt = F.linear(masked_mean(text_emb, text_mask), W["ptext.w"], W["ptext.b"])      # fp32 on both sides
        pp = F.linear(masked_mean(prompt_feats, prompt_mask), W["pprompt.w"], W["pprompt.b"])
        self.pooled = (pt + pp, pt + W["null_pp"][None, :])
        # per-layer cross-attention K/V for both branches
        self.kv = []
        for l in range(cfg.n_layers):
            w, b = W[f"l{l}.kv2.w"], W[f"l{l}.kv2.b"]
            self.kv.append(tuple(ops.lin(c, w, b, tag="kv2") for c in self.ctx))


def guidance_embedding(cfg, W, omega: float) -> torch.Tensor:
    """the guidance scale as an input of the distilled student — sinusoidal features of omega / 4
    (the time embedding's frequencies) through a two-layer MLP, added to the conditioning vector.  -> [d]"""
    feat = torch.tensor(S.time_features(omega / 4.0, cfg.d_time), dtype=torch.float64).to(torch.float32)
    return F.linear(F.silu(F.linear(feat, W["gs.w1"], W["gs.b1"])), W["gs.w2"], W["gs.b2"])


def denoiser_F(cfg, W, x_in: torch.Tensor, c_noise: float, cond: Conditioning, branch: int,
               ops: _Ops, g_emb=None) -> torch.Tensor:
    """a-4: F_theta(x_in, c_noise | text, prompt-or-null).  x_in [B,K,Ds] -> [B,K,Ds]."""
    d, L, H = cfg.d_model, cfg.n_layers, cfg.n_heads
    feat = torch.tensor(S.time_features(c_noise, cfg.d_time), dtype=torch.float64).to(torch.float32)
    t = F.linear(F.silu(F.linear(feat, W["time.w1"], W["time.b1"])), W["time.w2"], W["time.b2"])
    pre = t[None, :] + cond.pooled[branch]
    if g_emb is not None:
        pre = pre + g_emb[None, :]
    c = F.silu(pre)                                                     # [B,d]
    mod = ops.lin(c, W["mod.w"], W["mod.b"], tag="mod")                         # [B,(9L+2)d]

    def m(i):  # modulation chunk i -> [B,1,d]
        return mod[:, i * d:(i + 1) * d][:, None, :]

    h = ops.lin(x_in, W["in.w"], W["in.b"], tag="in") + W["pos"][None]
    for l in range(L):
        p, o = f"l{l}.", 9 * l
        u = layer_norm(h) * (1 + m(o + 1)) + m(o + 0)
        qkv = ops.lin(u, W[p + "qkv.w"], W[p + "qkv.b"], tag="qkv")
        a = attention(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], H, None, ops)
        h = h + m(o + 2) * ops.lin(a, W[p + "o.w"], W[p + "o.b"], tag="o")
        u = layer_norm(h) * (1 + m(o + 4)) + m(o + 3)
        q = ops.lin(u, W[p + "q2.w"], W[p + "q2.b"], tag="q2")
        kv = cond.kv[l][branch]
        a = attention(q, kv[..., :d], kv[..., d:], H, cond.ctx_mask[branch], ops)
        h = h + m(o + 5) * ops.lin(a, W[p + "o2.w"], W[p + "o2.b"], tag="o2")
        u = layer_norm(h) * (1 + m(o + 7)) + m(o + 6)
        f = gelu_tanh(ops.lin(u, W[p + "ff1.w"], W[p + "ff1.b"], tag="ff1"))
        h = h + m(o + 8) * ops.lin(f, W[p + "ff2.w"], W[p + "ff2.b"], tag="ff2")
    u = layer_norm(h) * (1 + m(9 * L + 1)) + m(9 * L)
    return ops.lin(u, W["out.w"], W["out.b"], tag="out")


def guided_denoise(cfg, W, x, sigma: float, cond: Conditioning, cfg_scale: float, ops: _Ops,
                   F_override=None):
    """a-2 + CFG (README.md:5 'classifier-free guidance'): D = c_skip x + c_out (F_u + w (F_c - F_u))."""
    c_skip, c_out, c_in, c_noise = S.edm_precond(sigma, cfg.sigma_data)
    if F_override is not None:
        Fg = F_override(c_in * x, sigma)
    else:
        Fc = denoiser_F(cfg, W, c_in * x, c_noise, cond, 0, ops)
        Fu = denoiser_F(cfg, W, c_in * x, c_noise, cond, 1, ops)
        Fg = Fu + cfg_scale * (Fc - Fu)
    return c_skip * x + c_out * Fg


def guided_student_denoise(cfg, W, x, sigma: float, cond: Conditioning, omega: float, ops: _Ops):
    """The guidance-conditioned student (README.md:5 "the style diffusion model is distilled"):
    D = c_skip x + c_out F(c_in x; sigma, omega | text, prompt) — ONE conditional branch, omega as an embedding."""
    c_skip, c_out, c_in, c_noise = S.edm_precond(sigma, cfg.sigma_data)
    Fg = denoiser_F(cfg, W, c_in * x, c_noise, cond, 0, ops, g_emb=guidance_embedding(cfg, W, omega))
    return c_skip * x + c_out * Fg


def sample_loop(cfg, denoise, noise: torch.Tensor, steps: int, sampler: int) -> torch.Tensor:
    """a-6/a-7: the sampler loop around D(x, sigma) = denoise(x, sigma)."""
    if sampler in (SAMPLER_STUDENT, SAMPLER_GUIDED):
        sig = S.student_sigmas(steps, cfg)
        x = sig[0] * noise[0]
        for i in range(steps):
            D = denoise(x, sig[i])
            d = (x - D) / sig[i]
            x = x + d * (sig[i + 1] - sig[i])
        return x
    sig = S.teacher_sigmas(steps, cfg)
    x = sig[0] * noise[0]
    for i in range(steps):
        s, sn = sig[i], sig[i + 1]
        s_up, s_down, s_mid = S.adpm2_sigmas(s, sn)
        d = (x - denoise(x, s)) / s
        x_mid = x + d * (s_mid - s)
        d_mid = (x_mid - denoise(x_mid, s_mid)) / s_mid
        x = x + d_mid * (s_down - s) + s_up * noise[i + 1]
    return x


class OraclePath:
    """Same module API as the CUDA path, CPU fp32."""

    def __init__(self, cfg, weights: torch.Tensor, emulate_bf16: bool = False, exact_tags=()):
        from styletts_zs_b200.spec import view_weights
        self.cfg = cfg
        self.W = view_weights(cfg, weights.detach().to(torch.float32).cpu())
        self.ops = _Ops(emulate_bf16, exact_tags)
        self._lstm = None

    # ---- a-7 ------------------------------------------------------------------------
    @torch.no_grad()
    def sample_style(self, text_emb, prompt_feats, steps: int, cfg_scale: float, *, text_mask=None,
                     prompt_mask=None, noise=None, sampler="student", seed: Optional[int] = None,
                     first_utterance=0) -> torch.Tensor:
        cfg = self.cfg
        B, T, _ = text_emb.shape
        P = prompt_feats.shape[1]
        kind = SAMPLER_TEACHER if sampler in ("teacher", SAMPLER_TEACHER) else \
            (SAMPLER_GUIDED if sampler in ("guided", SAMPLER_GUIDED) else SAMPLER_STUDENT)
        if noise is None and seed is not None:   # §8(f) rank 4: counter-based noise, oracle/philox.py
            from . import philox
            ns = steps + 1 if kind == SAMPLER_TEACHER else 1
            noise = torch.from_numpy(philox.normal_noise(seed, first_utterance, ns, B, cfg.n_style * cfg.d_style))
            noise = noise.reshape(ns, B, cfg.n_style, cfg.d_style)
        if text_mask is None:
            text_mask = torch.ones(B, T, dtype=torch.bool)
        if prompt_mask is None:
            prompt_mask = torch.ones(B, P, dtype=torch.bool)
        if noise is None:
            raise ValueError("noise tensor is an input (identical seeds == identical noise tensors)")
        cond = Conditioning(cfg, self.W, text_emb.float(), text_mask.bool(), prompt_feats.float(),
                            prompt_mask.bool(), self.ops)
        if kind == SAMPLER_GUIDED:
            den = lambda x, s: guided_student_denoise(cfg, self.W, x, s, cond, cfg_scale, self.ops)
        else:
            den = lambda x, s: guided_denoise(cfg, self.W, x, s, cond, cfg_scale, self.ops)
        return sample_loop(cfg, den, noise.float(), steps, kind)

    # ---- a-8 .. a-11 ----------------------------------------------------------------
    def _lstms(self):
        if self._lstm is None:
            self._lstm = [self._lstm_module(f"lstm{l}") for l in range(self.cfg.n_lstm)]
        return self._lstm

    @torch.no_grad()
    def regulate_length(self, feats, durations, *, max_frames=None, return_tokens=False):
        """Length regulator: torch.repeat_interleave per utterance, zero-padded to max_frames
        (default max_dur * T), totals truncated at max_frames.  -> (frames [B,F,C], frame_lens [B] int32[, frame_tok])."""
        feats, durations = feats.float(), durations.to(torch.int64).clamp(min=0)
        B, T, Cc = feats.shape
        F_max = int(max_frames) if max_frames is not None else self.cfg.max_dur * T
        frames = torch.zeros(B, F_max, Cc)
        tok = torch.full((B, F_max), -1, dtype=torch.int32)
        lens = torch.zeros(B, dtype=torch.int32)
        for b in range(B):
            idx = torch.repeat_interleave(torch.arange(T), durations[b])[:F_max]
            frames[b, :idx.numel()] = feats[b, idx]
            tok[b, :idx.numel()] = idx.to(torch.int32)
            lens[b] = idx.numel()
        return (frames, lens, tok) if return_tokens else (frames, lens)

    @torch.no_grad()
    def style_per_token(self, text_emb, style_codes):
        """a-8: s_tok = MHA(q = text, kv = style codes), n_sp_heads heads."""
        W, cfg = self.W, self.cfg
        q = F.linear(text_emb, W["sp.q.w"], W["sp.q.b"])
        k = F.linear(style_codes, W["sp.k.w"], W["sp.k.b"])
        v = F.linear(style_codes, W["sp.v.w"], W["sp.v.b"])
        a = attention(q, k, v, cfg.n_sp_heads, None, _Ops(False))
        return F.linear(a, W["sp.o.w"], W["sp.o.b"])

    def _lstm_module(self, prefix):
        cfg, W = self.cfg, self.W
        m = torch.nn.LSTM(cfg.d_hid + cfg.d_sty_tok, cfg.h_lstm, batch_first=True, bidirectional=True)
        with torch.no_grad():
            for dr, suf in (("f", ""), ("r", "_reverse")):
                p = f"{prefix}.{dr}."
                getattr(m, "weight_ih_l0" + suf).copy_(W[p + "w_ih"])
                getattr(m, "weight_hh_l0" + suf).copy_(W[p + "w_hh"])
                getattr(m, "bias_ih_l0" + suf).copy_(W[p + "b_ih"])
                getattr(m, "bias_hh_l0" + suf).copy_(W[p + "b_hh"])
        return m.eval()

    @staticmethod
    def _packed_bilstm(lstm, inp, lens, total_length):
        """BiLSTM with packed-sequence semantics (the reverse direction starts at each sequence's own last valid step);
        zero-length sequences produce zeros."""
        out = torch.zeros(inp.shape[0], total_length, 2 * lstm.hidden_size)
        nz = torch.nonzero(lens > 0).flatten()
        if nz.numel():
            packed = torch.nn.utils.rnn.pack_padded_sequence(inp[nz], lens[nz].cpu(), batch_first=True, enforce_sorted=False)
            o, _ = lstm(packed)
            o, _ = torch.nn.utils.rnn.pad_packed_sequence(o, batch_first=True, total_length=total_length)
            out[nz] = o
        return out

    @torch.no_grad()
    def _duration_encoder(self, text_emb, style_codes, text_mask):
        """a-8 + a-9: -> (d_enc [B,T,d_hid] = input of the final BiLSTM, x [B,T,d_hid] = its output, s_tok [B,T,d_sty_tok])."""
        cfg, W = self.cfg, self.W
        B, T, _ = text_emb.shape
        lens = text_mask.sum(1)
        # masks are prefix masks (padding at the end), as produced by a length vector
        assert bool((text_mask == (torch.arange(T)[None] < lens[:, None])).all()), "text_mask must be a prefix mask"
        mf = text_mask.to(torch.float32)[..., None]
        s_tok = self.style_per_token(text_emb, style_codes)
        x, d_enc = text_emb, None
        for l, lstm in enumerate(self._lstms()):
            if l == cfg.n_lstm - 1:
                d_enc = x
            inp = torch.cat([x, s_tok], -1)
            x = self._packed_bilstm(lstm, inp, lens, T)
            if l < cfg.n_lstm - 1:
                gb = F.linear(s_tok, W[f"adaln{l}.w"], W[f"adaln{l}.b"])
                x = layer_norm(x) * (1 + gb[..., :cfg.d_hid]) + gb[..., cfg.d_hid:]
                x = x * mf
        return d_enc, x, s_tok

    @torch.no_grad()
    def predict_duration(self, text_emb, style_codes, *, text_mask=None, return_presum=False):
        cfg, W = self.cfg, self.W
        text_emb, style_codes = text_emb.float(), style_codes.float()
        B, T, _ = text_emb.shape
        if text_mask is None:
            text_mask = torch.ones(B, T, dtype=torch.bool)
        text_mask = text_mask.bool()
        mf = text_mask.to(torch.float32)[..., None]
        _, x, _ = self._duration_encoder(text_emb, style_codes, text_mask)
        s = torch.sigmoid(F.linear(x, W["dur.w"], W["dur.b"])).sum(-1)           # a-10
        dur = (torch.round(s).clamp(min=1) * mf[..., 0]).to(torch.int32)        # half-to-even
        return (dur, s) if return_presum else dur

    @torch.no_grad()
    def predict_prosody(self, text_emb, style_codes, *, text_mask=None, durations=None, max_frames=None):
        """SURVEY.md §8(f) rank 2, second half — the F0 / energy ("prosody") heads behind the length regulator, pinned here
        (the reference has no code; the shape follows the StyleTTS lineage's F0Ntrain: a shared recurrent layer over the
        length-regulated duration-encoder features, then one small head per curve):
            en      = regulate([d_enc | s_tok], durations)                      [B,F,d_hid + d_sty_tok]
            y       = BiLSTM_pros(en)   (packed by frame length)                 [B,F,d_hid]
            z       = gelu_tanh([y | s_frame] W_h1^T + b_h1)                     [B,F,2 d_pros]   (s_frame = en[..., d_hid:])
            f0      = z[..., :d_pros] . w_f0 + b_f0,  energy = z[..., d_pros:] . w_en + b_en      0 past the frame length
        durations=None uses predict_duration's.  -> (f0 [B,F], energy [B,F], frame_lens [B] int32, durations [B,T] int32)."""
        cfg, W = self.cfg, self.W
        text_emb, style_codes = text_emb.float(), style_codes.float()
        B, T, _ = text_emb.shape
        if text_mask is None:
            text_mask = torch.ones(B, T, dtype=torch.bool)
        text_mask = text_mask.bool()
        mf = text_mask.to(torch.float32)[..., None]
        d_enc, x, s_tok = self._duration_encoder(text_emb, style_codes, text_mask)
        if durations is None:
            s = torch.sigmoid(F.linear(x, W["dur.w"], W["dur.b"])).sum(-1)
            durations = (torch.round(s).clamp(min=1) * mf[..., 0]).to(torch.int32)
        durations = durations.to(torch.int32)
        F_max = int(max_frames) if max_frames is not None else cfg.max_dur * T
        en, flen = self.regulate_length(torch.cat([d_enc, s_tok], -1), durations, max_frames=F_max)
        if not hasattr(self, "_pros_lstm"):
            self._pros_lstm = self._lstm_module("pros.lstm")
        y = self._packed_bilstm(self._pros_lstm, en, flen.to(torch.int64), F_max)
        z = gelu_tanh(F.linear(torch.cat([y, en[..., cfg.d_hid:]], -1), W["pros.h1.w"], W["pros.h1.b"]))
        fm = (torch.arange(F_max)[None] < flen[:, None]).to(torch.float32)
        f0 = (z[..., :cfg.d_pros] @ W["pros.f0.w"] + W["pros.f0.b"]) * fm
        en_curve = (z[..., cfg.d_pros:] @ W["pros.en.w"] + W["pros.en.b"]) * fm
        return f0, en_curve, flen, durations
