"""Pinned specification of the StyleTTS-ZS inference hot path.

The upstream repository ships no code (/root/reference/README.md:11-16 — both
"Training" and "Inference" are "Under construction"), so the contract between
the fp32 CPU oracle (``oracle/``) and the sm_100a CUDA path (``csrc/``) is this
file: dimensions, the flat fp32 weight-blob layout, the deterministic random
initialisation and the synthetic-input generator.  Everything that follows the
abstract's description of the path cites /root/reference/README.md:5.

The blob layout here is mirrored, entry for entry, by ``csrc/stz_layout.h``;
``tests/test_abi_cpu.py`` checks the two agree through ``stz_weight_offset``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Tuple

import torch

ABI_VERSION = 1

# sampler kinds (C-ABI enum stz_sampler_kind)
SAMPLER_STUDENT = 0   # distilled few-step student: deterministic Euler on a Karras grid + terminal 0
SAMPLER_TEACHER = 1   # undistilled teacher: ADPM2 (2 denoiser evals / step, ancestral noise)
SAMPLER_GUIDED = 2    # guidance-conditioned student: cfg_scale as an input embedding, ONE branch per step


@dataclass(frozen=True)
class StzConfig:
    """Dimensions of the path.  Field order == ``struct stz_config`` in include/stz.h."""
    n_style: int = 50        # K   fixed-length time-varying style codes (README.md:5)
    d_style: int = 512       # Ds  latent channels of one style code
    d_model: int = 512       # d   denoiser width
    n_heads: int = 8         # H
    d_ff: int = 2048
    n_layers: int = 8        # L
    d_text: int = 512
    d_prompt: int = 512
    d_time: int = 256        # sinusoidal features of c_noise
    # duration predictor
    d_hid: int = 512         # == d_text (x0 = text_emb)
    d_sty_tok: int = 128     # per-token style summary
    n_sp_heads: int = 4      # heads of the text->style pooling attention
    n_lstm: int = 4          # 3 x (BiLSTM + AdaLN) + 1 BiLSTM
    max_dur: int = 50
    # diffusion
    sigma_data: float = 0.5
    sigma_max: float = 3.0
    sigma_min: float = 1e-4
    rho: float = 9.0

    @property
    def d_head(self) -> int:
        return self.d_model // self.n_heads

    @property
    def n_mod(self) -> int:
        """AdaLN modulation vector length: 9 per layer (shift/scale/gate x 3) + 2 final."""
        return (9 * self.n_layers + 2) * self.d_model

    @property
    def h_lstm(self) -> int:
        return self.d_hid // 2

    @property
    def d_pros(self) -> int:
        """Hidden width of each prosody head (F0, energy)."""
        return self.d_hid // 2

    def as_dict(self):
        return asdict(self)


DEFAULT = StzConfig()
# A small configuration for fast CPU tests; keeps every structural feature.
TINY = StzConfig(n_style=10, d_style=64, d_model=64, n_heads=2, d_ff=128, n_layers=2,
                 d_text=64, d_prompt=64, d_time=32, d_hid=64, d_sty_tok=32,
                 n_sp_heads=2, n_lstm=4, max_dur=50)


# --------------------------------------------------------------------------------------
# weight blob layout
# --------------------------------------------------------------------------------------
def weight_entries(cfg: StzConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, init-kind) in blob order.  Linear weights are [out, in] (y = x W^T + b).

    init-kind: "w" fan-in scaled normal, "b" small normal, "e" embedding-like N(0, 0.5^2),
    "mw"/"mb" modulation weight/bias (non-zero: AdaLN-zero would make the net an identity and
    parity vacuous, SURVEY.md §7 hard part 6), "dw"/"db" duration-head weight/bias.
    """
    d, Ds, L = cfg.d_model, cfg.d_style, cfg.n_layers
    E: List[Tuple[str, Tuple[int, ...], str]] = []
    a = E.append
    # --- denoiser -------------------------------------------------------------------
    a(("in.w", (d, Ds), "w")); a(("in.b", (d,), "b"))
    a(("pos", (cfg.n_style, d), "e"))
    a(("time.w1", (d, cfg.d_time), "w")); a(("time.b1", (d,), "b"))
    a(("time.w2", (d, d), "w")); a(("time.b2", (d,), "b"))
    a(("ptext.w", (d, cfg.d_text), "w")); a(("ptext.b", (d,), "b"))
    a(("pprompt.w", (d, cfg.d_prompt), "w")); a(("pprompt.b", (d,), "b"))
    a(("null_pp", (d,), "e"))
    a(("ctx_text.w", (d, cfg.d_text), "w")); a(("ctx_text.b", (d,), "b"))
    a(("ctx_prompt.w", (d, cfg.d_prompt), "w")); a(("ctx_prompt.b", (d,), "b"))
    a(("type_emb", (2, d), "e"))
    a(("null_tok", (d,), "e"))
    a(("mod.w", (cfg.n_mod, d), "mw")); a(("mod.b", (cfg.n_mod,), "mb"))
    for l in range(L):
        p = f"l{l}."
        a((p + "qkv.w", (3 * d, d), "w")); a((p + "qkv.b", (3 * d,), "b"))
        a((p + "o.w", (d, d), "w")); a((p + "o.b", (d,), "b"))
        a((p + "q2.w", (d, d), "w")); a((p + "q2.b", (d,), "b"))
        a((p + "kv2.w", (2 * d, d), "w")); a((p + "kv2.b", (2 * d,), "b"))
        a((p + "o2.w", (d, d), "w")); a((p + "o2.b", (d,), "b"))
        a((p + "ff1.w", (cfg.d_ff, d), "w")); a((p + "ff1.b", (cfg.d_ff,), "b"))
        a((p + "ff2.w", (d, cfg.d_ff), "w")); a((p + "ff2.b", (d,), "b"))
    a(("out.w", (Ds, d), "w")); a(("out.b", (Ds,), "b"))
    # --- duration predictor ---------------------------------------------------------
    ds, dh, h = cfg.d_sty_tok, cfg.d_hid, cfg.h_lstm
    a(("sp.q.w", (ds, cfg.d_text), "w")); a(("sp.q.b", (ds,), "b"))
    a(("sp.k.w", (ds, Ds), "w")); a(("sp.k.b", (ds,), "b"))
    a(("sp.v.w", (ds, Ds), "w")); a(("sp.v.b", (ds,), "b"))
    a(("sp.o.w", (ds, ds), "w")); a(("sp.o.b", (ds,), "b"))
    for l in range(cfg.n_lstm):
        for dr in ("f", "r"):
            p = f"lstm{l}.{dr}."
            a((p + "w_ih", (4 * h, dh + ds), "w")); a((p + "w_hh", (4 * h, h), "w"))
            a((p + "b_ih", (4 * h,), "b")); a((p + "b_hh", (4 * h,), "b"))
        if l < cfg.n_lstm - 1:
            a((f"adaln{l}.w", (2 * dh, ds), "w")); a((f"adaln{l}.b", (2 * dh,), "b"))
    a(("dur.w", (cfg.max_dur, dh), "dw")); a(("dur.b", (cfg.max_dur,), "db"))
    # --- prosody (F0 / energy) heads behind the length regulator -----------------------------
    # appended after every entry above, so the offsets (and the seeded values) of the path's weights do not move
    for dr in ("f", "r"):
        p = f"pros.lstm.{dr}."
        a((p + "w_ih", (4 * h, dh + ds), "w")); a((p + "w_hh", (4 * h, h), "w"))
        a((p + "b_ih", (4 * h,), "b")); a((p + "b_hh", (4 * h,), "b"))
    a(("pros.h1.w", (2 * cfg.d_pros, dh + ds), "w")); a(("pros.h1.b", (2 * cfg.d_pros,), "b"))
    a(("pros.f0.w", (cfg.d_pros,), "w")); a(("pros.f0.b", (1,), "b"))
    a(("pros.en.w", (cfg.d_pros,), "w")); a(("pros.en.b", (1,), "b"))
    # --- guidance-scale embedding of the guidance-conditioned student, appended ---------------
    a(("gs.w1", (d, cfg.d_time), "w")); a(("gs.b1", (d,), "b"))
    a(("gs.w2", (d, d), "w")); a(("gs.b2", (d,), "b"))
    return E


def weight_offsets(cfg: StzConfig) -> Dict[str, Tuple[int, Tuple[int, ...]]]:
    """name -> (offset in floats, shape).  Every entry is padded to a multiple of 64 floats
    (256 B) so device-side views are 256-byte aligned for TMA and 128-bit loads."""
    off, out = 0, {}
    for name, shape, _ in weight_entries(cfg):
        n = math.prod(shape)
        out[name] = (off, shape)
        off += (n + 63) // 64 * 64
    out["__total__"] = (off, ())
    return out


def init_weights(cfg: StzConfig = DEFAULT, seed: int = 0) -> torch.Tensor:
    """Deterministic random-init flat fp32 blob (north_star: 'random-init weights')."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    offs = weight_offsets(cfg)
    blob = torch.zeros(offs["__total__"][0], dtype=torch.float32)
    for name, shape, kind in weight_entries(cfg):
        off, _ = offs[name]
        n = math.prod(shape)
        r = torch.randn(n, generator=g, dtype=torch.float32)
        if kind == "w":
            r *= 1.0 / math.sqrt(shape[-1])
        elif kind == "b":
            r *= 0.02
        elif kind == "e":
            r *= 0.5
        elif kind == "mw":
            r *= 0.5 / math.sqrt(shape[-1])
        elif kind == "mb":
            r *= 0.02
        elif kind == "dw":   # wide logits -> per-token durations spread over the 1..max_dur range
            r *= 4.0 / math.sqrt(shape[-1])
        elif kind == "db":
            r = r * 1.0 - 1.5
        else:  # pragma: no cover
            raise ValueError(kind)
        blob[off:off + n] = r
    # gates: bias them to +-0.5 around a non-zero mean so residual branches are active
    d, L = cfg.d_model, cfg.n_layers
    off, _ = offs["mod.b"]
    for l in range(L):
        for s in range(3):
            lo = off + (l * 9 + s * 3 + 2) * d
            blob[lo:lo + d] += 0.5
    return blob


def view_weights(cfg: StzConfig, blob: torch.Tensor) -> Dict[str, torch.Tensor]:
    offs = weight_offsets(cfg)
    assert blob.numel() == offs["__total__"][0], (blob.numel(), offs["__total__"][0])
    out = {}
    for name, (off, shape) in offs.items():
        if name == "__total__":
            continue
        out[name] = blob[off:off + math.prod(shape)].view(*shape)
    return out


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def n_noise_slices(steps: int, sampler: int) -> int:
    """Slice 0 is the initial noise; teacher step i (1-based) reads slice i."""
    return steps + 1 if sampler == SAMPLER_TEACHER else 1


def synthetic_inputs(cfg: StzConfig, B: int, T: int, *, P: int | None = None, steps: int = 4,
                     sampler: int = SAMPLER_STUDENT, seed: int = 1234,
                     var_len: Tuple[int, int] | None = None):
    """text_emb [B,T,d_text], text_mask [B,T] bool (True = valid), prompt_feats [B,P,d_prompt],
    prompt_mask [B,P], noise [n_slices,B,K,Ds]; all fp32 on CPU from one seeded generator."""
    P = cfg.n_style if P is None else P
    g = torch.Generator(device="cpu").manual_seed(seed)
    text = torch.randn(B, T, cfg.d_text, generator=g)
    prompt = torch.randn(B, P, cfg.d_prompt, generator=g)
    noise = torch.randn(n_noise_slices(steps, sampler), B, cfg.n_style, cfg.d_style, generator=g)
    if var_len is None:
        lens = torch.full((B,), T, dtype=torch.int64)
    else:
        lo, hi = var_len
        lens = torch.randint(lo, min(hi, T) + 1, (B,), generator=g)
    text_mask = torch.arange(T)[None, :] < lens[:, None]
    prompt_mask = torch.ones(B, P, dtype=torch.bool)
    return dict(text_emb=text, text_mask=text_mask, prompt_feats=prompt, prompt_mask=prompt_mask,
                noise=noise, lens=lens)
