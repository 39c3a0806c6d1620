"""Host side of the drop-in boundary: a ctypes binding of include/stz.h behind the module API
``sample_style(text_emb, prompt_feats, steps, cfg_scale)`` / ``predict_duration(...)``
(BASELINE.json north_star).  The reference has no such module to cite
(/root/reference/README.md:15-16); the oracle in ``oracle/`` exposes the same signatures so the
parity tests read identically for both.

PyTorch is plumbing only here: device memory, streams, (for multi-GPU) process launch.
There is NO CPU fallback and nothing in this module imports ``oracle``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch

from .spec import (ABI_VERSION, DEFAULT, SAMPLER_GUIDED, SAMPLER_STUDENT, SAMPLER_TEACHER, StzConfig, n_noise_slices,
                   weight_offsets)

# STZ_LIBRARY: an alternative build of the same sources (the timeline build libstz_trace.so of tools/*_trace.py)
_LIB_PATH = os.environ.get("STZ_LIBRARY") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libstz.so")
_lib = None


class StzError(RuntimeError):
    pass


class _CConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_style", "d_style", "d_model", "n_heads", "d_ff", "n_layers",
                                          "d_text", "d_prompt", "d_time", "d_hid", "d_sty_tok",
                                          "n_sp_heads", "n_lstm", "max_dur")] + \
               [(n, C.c_float) for n in ("sigma_data", "sigma_max", "sigma_min", "rho")]


def _cconfig(cfg: StzConfig) -> _CConfig:
    return _CConfig(**{k: v for k, v in cfg.as_dict().items()})


def load_library(path: Optional[str] = None):
    """dlopen csrc/libstz.so and declare every prototype of include/stz.h.  Raises if missing:
    the product path must fail loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or _LIB_PATH
    if not os.path.exists(p):
        raise StzError(f"{p} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       f"(there is no CPU fallback)")
    lib = C.CDLL(p)
    vp, i32, f32, i64, sz = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_size_t
    cfgp = C.POINTER(_CConfig)
    lib.stz_abi_version.restype = i32
    lib.stz_abi_version.argtypes = []
    lib.stz_weights_nfloats.restype = sz
    lib.stz_weights_nfloats.argtypes = [cfgp]
    lib.stz_weight_offset.restype = i64
    lib.stz_weight_offset.argtypes = [cfgp, C.c_char_p]
    lib.stz_create.restype = i32
    lib.stz_create.argtypes = [cfgp, vp, sz, i32, C.POINTER(vp)]
    lib.stz_destroy.restype = None
    lib.stz_destroy.argtypes = [vp]
    lib.stz_last_error.restype = C.c_char_p
    lib.stz_last_error.argtypes = [vp]
    lib.stz_sample_style.restype = i32
    lib.stz_sample_style.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, i32, vp, vp]
    lib.stz_predict_duration.restype = i32
    lib.stz_predict_duration.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp]
    lib.stz_regulate_length.restype = i32
    lib.stz_regulate_length.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.stz_predict_prosody.restype = i32
    lib.stz_predict_prosody.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.stz_synthesize_host.restype = i32
    lib.stz_synthesize_host.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, i32, vp, vp]
    lib.stz_synthesize_host_submit.restype = i32
    lib.stz_synthesize_host_submit.argtypes = [vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, i32, vp, vp]
    lib.stz_synthesize_host_wait.restype = i32
    lib.stz_synthesize_host_wait.argtypes = [vp, i32]
    lib.stz_set_noise_seed.restype = i32
    lib.stz_set_noise_seed.argtypes = [vp, C.c_uint64, C.c_uint64]
    lib.stz_set_noise_utterances.restype = i32
    lib.stz_set_noise_utterances.argtypes = [vp, C.c_uint64, C.POINTER(C.c_uint64), i32]
    lib.stz_philox_normal.restype = i32
    lib.stz_philox_normal.argtypes = [C.c_uint64, C.c_uint64, i32, i32, i32, vp, i32, vp]
    lib.stz_debug_plan.restype = i32
    lib.stz_debug_plan.argtypes = [cfgp, i32, i32, f32, vp, vp, vp, vp]
    lib.stz_launch_count.restype = i64
    lib.stz_launch_count.argtypes = [vp]
    lib.stz_set_option.restype = i32
    lib.stz_set_option.argtypes = [vp, C.c_char_p, i32]
    lib.stz_get_option.restype = i32
    lib.stz_get_option.argtypes = [vp, C.c_char_p, C.POINTER(i32)]
    lib.stz_debug_check_guards.restype = i32
    lib.stz_debug_check_guards.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.stz_profile_read.restype = i32
    lib.stz_profile_read.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64)]
    lib.stz_bench_gemm.restype = i32
    lib.stz_bench_gemm.argtypes = [vp, i32, i32, i32, i32, i32, C.POINTER(C.c_double)]
    lib.stz_debug_max_lstm_clusters.restype = i32
    lib.stz_debug_max_lstm_clusters.argtypes = []
    lib.stz_debug_set_tap.restype = i32
    lib.stz_debug_set_tap.argtypes = [vp, i32, i32, i32, vp]
    lib.stz_debug_set_lstm_trace.restype = i32
    lib.stz_debug_set_lstm_trace.argtypes = [vp, vp]
    lib.stz_debug_set_gemm_trace.restype = i32
    lib.stz_debug_set_gemm_trace.argtypes = [vp, vp]
    lib.stz_debug_set_att_trace.restype = i32
    lib.stz_debug_set_att_trace.argtypes = [vp, vp]
    lib.stz_op_attention.restype = i32
    lib.stz_op_attention.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, vp, i32, vp]
    lib.stz_op_gemm_bf16.restype = i32
    lib.stz_op_gemm_bf16.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.stz_op_gemm_epi.restype = i32
    lib.stz_op_gemm_epi.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, i32, i32, vp, vp]
    lib.stz_op_gemm_sampler.restype = i32
    lib.stz_op_gemm_sampler.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.stz_op_gemm_ln.restype = i32
    lib.stz_op_gemm_ln.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, i32, i32, i32, i32, vp, i32, vp, vp]
    lib.stz_graph_count.restype = i32
    lib.stz_graph_count.argtypes = [vp, C.POINTER(i64)]
    lib.stz_reserve.restype = i32
    lib.stz_reserve.argtypes = [vp, i32, i32, i32, i32, i32]
    if lib.stz_abi_version() != ABI_VERSION:
        raise StzError(f"ABI mismatch: library {lib.stz_abi_version()} vs python {ABI_VERSION}")
    if path is None:
        _lib = lib
    return lib


EXPORTED_SYMBOLS = ("stz_abi_version", "stz_weights_nfloats", "stz_weight_offset", "stz_create",
                    "stz_destroy", "stz_last_error", "stz_sample_style", "stz_predict_duration",
                    "stz_synthesize_host", "stz_synthesize_host_submit", "stz_synthesize_host_wait", "stz_regulate_length", "stz_set_noise_seed", "stz_set_noise_utterances", "stz_philox_normal", "stz_debug_plan", "stz_launch_count", "stz_set_option", "stz_profile_read",
                    "stz_debug_set_tap", "stz_debug_set_att_trace", "stz_debug_set_gemm_trace", "stz_debug_set_lstm_trace", "stz_debug_max_lstm_clusters", "stz_bench_gemm",
                    "stz_op_gemm_bf16", "stz_op_attention", "stz_op_gemm_epi", "stz_op_gemm_sampler", "stz_op_gemm_ln",
                    "stz_graph_count", "stz_reserve", "stz_get_option", "stz_predict_prosody", "stz_debug_check_guards")


def _kind(sampler) -> int:
    if sampler in ("student", SAMPLER_STUDENT):
        return SAMPLER_STUDENT
    if sampler in ("teacher", SAMPLER_TEACHER):
        return SAMPLER_TEACHER
    if sampler in ("guided", SAMPLER_GUIDED):
        return SAMPLER_GUIDED
    raise ValueError(f"unknown sampler {sampler!r}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class StyleTTSZSPath:
    """CUDA (sm_100a) implementation of the hot path behind the oracle's module API."""

    def __init__(self, cfg: StzConfig = DEFAULT, weights: Optional[torch.Tensor] = None, device: int = 0,
                 backend: str = "cuda"):
        if backend != "cuda":
            raise StzError("only backend='cuda' exists in the product package; the fp32 CPU oracle lives "
                           "in oracle/ and is test infrastructure")
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise StzError("no CUDA device: the hot path has no CPU fallback")
        self.cfg = cfg
        self.device = torch.device("cuda", device)
        if weights is None:
            from .spec import init_weights
            weights = init_weights(cfg, 0)
        w = weights.detach().to(torch.float32).cpu().contiguous()
        n = weight_offsets(cfg)["__total__"][0]
        if w.numel() != n or self.lib.stz_weights_nfloats(C.byref(_cconfig(cfg))) != n:
            raise StzError("weight blob size does not match the layout")
        h = C.c_void_p()
        rc = self.lib.stz_create(C.byref(_cconfig(cfg)), C.c_void_p(w.data_ptr()), n, device, C.byref(h))
        if rc != 0:
            raise StzError(f"stz_create failed ({rc}): {self.lib.stz_last_error(None).decode()}")
        self._h = h
        self._inflight = {}

    def close(self):
        if getattr(self, "_h", None):
            self.lib.stz_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise StzError(f"{what} failed ({rc}): {self.lib.stz_last_error(self._h).decode()}")

    # ---------------------------------------------------------------------------------
    def set_option(self, key: str, value: int):
        self._check(self.lib.stz_set_option(self._h, key.encode(), int(value)), f"set_option({key})")

    def get_option(self, key: str) -> int:
        v = C.c_int()
        self._check(self.lib.stz_get_option(self._h, key.encode(), C.byref(v)), f"get_option({key})")
        return int(v.value)

    def check_guards(self) -> Tuple[int, int]:
        """(gaps checked, overwritten guard bytes) — see include/stz.h: stz_debug_check_guards; needs set_option('guard_bytes', n)."""
        bad = C.c_longlong()
        n = self.lib.stz_debug_check_guards(self._h, C.byref(bad))
        if n < 0:
            raise StzError(f"stz_debug_check_guards failed ({n}): {self.lib.stz_last_error(self._h).decode()}")
        return int(n), int(bad.value)

    def launch_count(self) -> int:
        return int(self.lib.stz_launch_count(self._h))

    def graph_count(self) -> Tuple[int, int]:
        """(cached CUDA graphs, captures since creation) of the evaluation loop (include/stz.h: stz_graph_count)."""
        tot = C.c_int64()
        n = self.lib.stz_graph_count(self._h, C.byref(tot))
        if n < 0:
            raise StzError("stz_graph_count failed")
        return int(n), int(tot.value)

    def reserve(self, max_B: int, max_T: int, max_P: Optional[int] = None, max_steps: int = 4, sampler="student"):
        """Size the workspace once for the largest call (no reallocation -> captured graphs survive)."""
        P = self.cfg.n_style if max_P is None else max_P
        self._check(self.lib.stz_reserve(self._h, int(max_B), int(max_T), int(P), int(max_steps), _kind(sampler)),
                    "stz_reserve")

    # ---- unit-test entries: fused epilogues of the product GEMM kernels ------------------------------------------------
    def op_gemm_epi(self, A, W, bias, epi: int, *, out=None, mod=None, gate_off: int = 0, pos=None) -> torch.Tensor:
        M, K = A.shape
        N = W.shape[0]
        if out is None:
            out = torch.empty(M, N, device=A.device, dtype=torch.bfloat16 if epi in (2, 3) else torch.float32)
        st = torch.cuda.current_stream().cuda_stream
        n_mod = 0 if mod is None else mod.shape[1]
        self._check(self.lib.stz_op_gemm_epi(self._h, _ptr(A), _ptr(W), _ptr(bias), M, N, K, int(epi), _ptr(out), _ptr(mod),
                                             n_mod, int(gate_off), _ptr(pos), C.c_void_p(st)), "stz_op_gemm_epi")
        return out

    def op_gemm_sampler(self, A, W, bias, x, xmid, noise, coef, *, tap=None) -> torch.Tensor:
        """Updates x / xmid in place, returns xin [M, 3N] bf16."""
        M, K = A.shape
        N = W.shape[0]
        xin = torch.empty(M, 3 * N, device=A.device, dtype=torch.bfloat16)
        st = torch.cuda.current_stream().cuda_stream
        self._check(self.lib.stz_op_gemm_sampler(self._h, _ptr(A), _ptr(W), _ptr(bias), M, N, K, _ptr(x), _ptr(xmid),
                                                 _ptr(noise), _ptr(coef), _ptr(xin), _ptr(tap), C.c_void_p(st)),
                    "stz_op_gemm_sampler")
        return xin

    def op_gemm_ln(self, A, W, bias, h, mod, *, mode: int = 0, gate_off: int = 0, shift_off: int = 0, scale_off: int = 0,
                   pos=None, split3: bool = False) -> torch.Tensor:
        """Overwrites h with h', returns u [M, 512] (or [M, 1536] if split3) bf16."""
        M, K = A.shape
        u = torch.empty(M, (3 if split3 else 1) * self.cfg.d_model, device=A.device, dtype=torch.bfloat16)
        st = torch.cuda.current_stream().cuda_stream
        self._check(self.lib.stz_op_gemm_ln(self._h, _ptr(A), _ptr(W), _ptr(bias), M, K, int(mode), _ptr(h), _ptr(mod),
                                            mod.shape[1], int(gate_off), int(shift_off), int(scale_off), _ptr(pos),
                                            1 if split3 else 0, _ptr(u), C.c_void_p(st)), "stz_op_gemm_ln")
        return u

    PROFILE_CLASSES = ("gemm_tc", "attention", "ln_mod", "linear_f32", "lstm_rec", "pred_ew", "other")

    def profile_read(self):
        """{class: (ms, work, launches)} accumulated since set_option('profile', 1)."""
        out = {}
        for i, name in enumerate(self.PROFILE_CLASSES):
            ms, work, n = C.c_double(), C.c_double(), C.c_int64()
            self._check(self.lib.stz_profile_read(self._h, i, C.byref(ms), C.byref(work), C.byref(n)), "profile_read")
            out[name] = (ms.value, work.value, n.value)
        return out

    def bench_gemm(self, M: int, N: int, K: int, epi: int, iters: int = 50) -> float:
        """Mean microseconds per launch of the product GEMM kernel for one shape (back-to-back launches)."""
        us = C.c_double()
        self._check(self.lib.stz_bench_gemm(self._h, M, N, K, epi, iters, C.byref(us)), "stz_bench_gemm")
        return us.value

    def op_attention(self, qkv: torch.Tensor, kv_text=None, kv_prompt=None, kv_null=None, text_mask=None,
                     prompt_mask=None, impl: int = 0) -> torch.Tensor:
        """Unit-test entry: the fused attention kernels on bf16 CUDA buffers (see include/stz.h: stz_op_attention)."""
        d, K = self.cfg.d_model, self.cfg.n_style
        R, ldq = qkv.shape
        B = R // (2 * K)
        T = 0 if kv_text is None else kv_text.shape[0] // B
        P = 0 if kv_prompt is None else kv_prompt.shape[0] // B
        out = torch.empty(R, d, dtype=torch.bfloat16, device=qkv.device)
        tm, pm = self._mask(text_mask), self._mask(prompt_mask)     # kept alive until the launch is enqueued
        st = torch.cuda.current_stream().cuda_stream
        self._check(self.lib.stz_op_attention(self._h, _ptr(qkv), ldq, _ptr(kv_text), _ptr(kv_prompt), _ptr(kv_null),
                                              _ptr(tm), _ptr(pm), B, T, P, _ptr(out), impl, C.c_void_p(st)),
                    "stz_op_attention")
        for t in (tm, pm):
            if t is not None:
                t.record_stream(torch.cuda.current_stream())
        return out

    def set_tap(self, ev: int, layer: int, stage: int, buf: Optional[torch.Tensor]):
        self._check(self.lib.stz_debug_set_tap(self._h, ev, layer, stage, _ptr(buf)), "set_tap")

    def _dev(self, t: torch.Tensor, dtype) -> torch.Tensor:
        return t.to(device=self.device, dtype=dtype, non_blocking=True).contiguous()

    def _mask(self, m: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        return None if m is None else self._dev(m.to(torch.uint8) if m.dtype == torch.bool else m, torch.uint8)

    # ---------------------------------------------------------------------------------
    def sample_style(self, text_emb, prompt_feats, steps: int, cfg_scale: float, *, text_mask=None,
                     prompt_mask=None, noise=None, sampler="student", seed: Optional[int] = None,
                     first_utterance=0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """-> style codes [B,K,Ds] fp32 on this path's device.  ``noise`` [n_slices,B,K,Ds] is an
        input ("identical seeds" == identical noise tensors); or pass ``seed``
        (and the global index of the batch's first utterance) to draw it on the device — the same
        tensor as ``philox_normal(seed, first_utterance, n_slices, B)``."""
        cfg, kind = self.cfg, _kind(sampler)
        B, T, _ = text_emb.shape
        P = prompt_feats.shape[1]
        if (noise is None) == (seed is None):
            raise ValueError("pass exactly one of noise= (tensor) or seed= (on-device Philox noise)")
        ns = n_noise_slices(steps, kind)
        if noise is not None and tuple(noise.shape) != (ns, B, cfg.n_style, cfg.d_style):
            raise ValueError(f"noise must be {(ns, B, cfg.n_style, cfg.d_style)}, got {tuple(noise.shape)}")
        if seed is not None:
            self.seed_noise(seed, first_utterance)
        with torch.cuda.device(self.device):
            te, pf = self._dev(text_emb, torch.float32), self._dev(prompt_feats, torch.float32)
            nz = None if noise is None else self._dev(noise, torch.float32)
            tm, pm = self._mask(text_mask), self._mask(prompt_mask)
            if out is None:
                out = torch.empty(B, cfg.n_style, cfg.d_style, dtype=torch.float32, device=self.device)
            elif tuple(out.shape) != (B, cfg.n_style, cfg.d_style) or out.dtype != torch.float32 or not out.is_contiguous():
                raise ValueError("out must be a contiguous fp32 CUDA tensor [B, K, Ds]")
            st = torch.cuda.current_stream().cuda_stream
            rc = self.lib.stz_sample_style(self._h, _ptr(te), _ptr(tm), _ptr(pf), _ptr(pm), _ptr(nz), B, T, P,
                                           int(steps), float(cfg_scale), kind, _ptr(out), C.c_void_p(st))
            self._check(rc, "stz_sample_style")
            # inputs converted above are kept alive until the work is enqueued *and* ordered:
            for t in (te, pf, nz, tm, pm):
                if t is not None:
                    t.record_stream(torch.cuda.current_stream())
        return out

    def seed_noise(self, seed: int, first_utterance=0):
        """Calls that pass no noise tensor draw it on the device from (seed, global utterance index).
        ``first_utterance``: an int (utterance b is first_utterance + b) or a sequence of B explicit indices."""
        if isinstance(first_utterance, int):
            self._check(self.lib.stz_set_noise_seed(self._h, int(seed) & (2 ** 64 - 1), int(first_utterance)),
                        "stz_set_noise_seed")
        else:
            ids = [int(i) for i in first_utterance]
            arr = (C.c_uint64 * len(ids))(*ids)
            self._check(self.lib.stz_set_noise_utterances(self._h, int(seed) & (2 ** 64 - 1), arr, len(ids)),
                        "stz_set_noise_utterances")

    def philox_normal(self, seed: int, first_utterance: int, slices: int, B: int) -> torch.Tensor:
        """The on-device generator's output [slices, B, K, Ds] fp32 (bit-identical to the oracle's)."""
        cfg = self.cfg
        out = torch.empty(slices, B, cfg.n_style, cfg.d_style, dtype=torch.float32, device=self.device)
        return philox_normal(seed, first_utterance, slices, B, cfg.n_style * cfg.d_style, out=out)

    def predict_duration(self, text_emb, style_codes, *, text_mask=None, return_presum=False):
        """-> int32 frames per token [B,T] (0 on padding)."""
        B, T, _ = text_emb.shape
        with torch.cuda.device(self.device):
            te, sc = self._dev(text_emb, torch.float32), self._dev(style_codes, torch.float32)
            tm = self._mask(text_mask)
            out = torch.empty(B, T, dtype=torch.int32, device=self.device)
            pre = torch.empty(B, T, dtype=torch.float32, device=self.device) if return_presum else None
            st = torch.cuda.current_stream().cuda_stream
            rc = self.lib.stz_predict_duration(self._h, _ptr(te), _ptr(tm), _ptr(sc), B, T, _ptr(out), _ptr(pre),
                                               C.c_void_p(st))
            self._check(rc, "stz_predict_duration")
            for t in (te, sc, tm):
                if t is not None:
                    t.record_stream(torch.cuda.current_stream())
        return (out, pre) if return_presum else out

    def predict_prosody(self, text_emb, style_codes, *, text_mask=None, durations=None, max_frames: Optional[int] = None):
        """F0 / energy heads behind the length regulator -> (f0 [B,F], energy [B,F] fp32, frame_lens [B] int32,
        durations [B,T] int32 as predicted); ``durations`` overrides the predicted ones for the regulator."""
        B, T, _ = text_emb.shape
        F_max = int(max_frames) if max_frames is not None else self.cfg.max_dur * T
        with torch.cuda.device(self.device):
            te, sc = self._dev(text_emb, torch.float32), self._dev(style_codes, torch.float32)
            tm = self._mask(text_mask)
            du = None if durations is None else self._dev(durations, torch.int32)
            f0 = torch.empty(B, F_max, dtype=torch.float32, device=self.device)
            en = torch.empty(B, F_max, dtype=torch.float32, device=self.device)
            fl = torch.empty(B, dtype=torch.int32, device=self.device)
            dur = torch.empty(B, T, dtype=torch.int32, device=self.device)
            st = torch.cuda.current_stream().cuda_stream
            rc = self.lib.stz_predict_prosody(self._h, _ptr(te), _ptr(tm), _ptr(sc), _ptr(du), B, T, F_max, _ptr(f0), _ptr(en),
                                              _ptr(fl), _ptr(dur), C.c_void_p(st))
            self._check(rc, "stz_predict_prosody")
            for t in (te, sc, tm, du):
                if t is not None:
                    t.record_stream(torch.cuda.current_stream())
        return f0, en, fl, dur

    def regulate_length(self, feats, durations, *, max_frames: Optional[int] = None, return_tokens: bool = False):
        """Length regulator: frames[b, f] = feats[b, token of frame f] for the integer durations of predict_duration.
        -> (frames [B,F,C] fp32, frame_lens [B] int32[, frame_tok [B,F] int32]); F = max_frames or max_dur * T."""
        B, T, Cc = feats.shape
        F_max = int(max_frames) if max_frames is not None else self.cfg.max_dur * T
        with torch.cuda.device(self.device):
            ft, du = self._dev(feats, torch.float32), self._dev(durations, torch.int32)
            frames = torch.empty(B, F_max, Cc, dtype=torch.float32, device=self.device)
            lens = torch.empty(B, dtype=torch.int32, device=self.device)
            tok = torch.empty(B, F_max, dtype=torch.int32, device=self.device) if return_tokens else None
            st = torch.cuda.current_stream().cuda_stream
            rc = self.lib.stz_regulate_length(self._h, _ptr(ft), _ptr(du), B, T, Cc, F_max, _ptr(frames), _ptr(lens),
                                              _ptr(tok), C.c_void_p(st))
            self._check(rc, "stz_regulate_length")
            for t in (ft, du):
                t.record_stream(torch.cuda.current_stream())
        return (frames, lens, tok) if return_tokens else (frames, lens)

    def synthesize_host(self, text_emb, prompt_feats, steps: int, cfg_scale: float, *, text_mask=None,
                        prompt_mask=None, noise=None, sampler="student", out_style=None, out_dur=None,
                        with_duration=True, seed: Optional[int] = None, first_utterance=0,
                        slot: Optional[int] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """End-to-end call with HOST tensors (pinned recommended): H2D + sample_style
        [+ predict_duration] + D2H + sync inside one C-ABI call.
        With ``seed`` instead of ``noise`` the noise is drawn on the device (no noise H2D).
        With ``slot`` (0 or 1) the call only SUBMITS the batch (stz_synthesize_host_submit) and returns the output
        tensors immediately; they are valid after ``synthesize_host_wait(slot)``.  Alternating the two slots overlaps
        batch i+1's input copies with batch i's compute — this is what bench.py's `e2e` times."""
        cfg, kind = self.cfg, _kind(sampler)
        B, T, _ = text_emb.shape
        P = prompt_feats.shape[1]
        if (noise is None) == (seed is None):
            raise ValueError("pass exactly one of noise= (tensor) or seed= (on-device Philox noise)")
        if seed is not None:
            self.seed_noise(seed, first_utterance)
        for t in (text_emb, prompt_feats) + (() if noise is None else (noise,)):
            if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("synthesize_host takes contiguous fp32 CPU tensors")
        tm = None if text_mask is None else text_mask.to(torch.uint8).contiguous()
        pm = None if prompt_mask is None else prompt_mask.to(torch.uint8).contiguous()
        if out_style is None:
            out_style = torch.empty(B, cfg.n_style, cfg.d_style, dtype=torch.float32)
        if with_duration and out_dur is None:
            out_dur = torch.empty(B, T, dtype=torch.int32)
        if slot is None:
            rc = self.lib.stz_synthesize_host(self._h, _ptr(text_emb), _ptr(tm), _ptr(prompt_feats), _ptr(pm),
                                              _ptr(noise), B, T, P, int(steps), float(cfg_scale), kind,
                                              _ptr(out_style), _ptr(out_dur) if with_duration else None)
            self._check(rc, "stz_synthesize_host")
        else:
            rc = self.lib.stz_synthesize_host_submit(self._h, int(slot), _ptr(text_emb), _ptr(tm), _ptr(prompt_feats),
                                                     _ptr(pm), _ptr(noise), B, T, P, int(steps), float(cfg_scale), kind,
                                                     _ptr(out_style), _ptr(out_dur) if with_duration else None)
            self._check(rc, "stz_synthesize_host_submit")
            # host buffers belong to the library until the slot's wait: keep them (and converted masks) alive
            self._inflight[int(slot)] = (text_emb, tm, prompt_feats, pm, noise, out_style, out_dur)
        return out_style, (out_dur if with_duration else None)

    def synthesize_host_wait(self, slot: int):
        """Blocks until the batch submitted into ``slot`` has its outputs in the host tensors."""
        self._check(self.lib.stz_synthesize_host_wait(self._h, int(slot)), "stz_synthesize_host_wait")
        self._inflight.pop(int(slot), None)


def op_gemm_bf16(A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], impl: int = 0) -> torch.Tensor:
    """Unit-test entry: C = A·W^T + bias with the library's GEMM (A [M,K], W [N,K] bf16 CUDA)."""
    lib = load_library()
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty(M, N, dtype=torch.float32, device=A.device)
    st = torch.cuda.current_stream(A.device).cuda_stream
    rc = lib.stz_op_gemm_bf16(_ptr(A), _ptr(W), _ptr(bias), _ptr(out), M, N, K, impl, A.device.index or 0,
                              C.c_void_p(st))
    if rc != 0:
        raise StzError(f"stz_op_gemm_bf16 failed ({rc}): {lib.stz_last_error(None).decode()}")
    return out


def philox_normal(seed: int, first_utterance: int, slices: int, B: int, n_per_utt: int,
                  out: Optional[torch.Tensor] = None, device: int = 0) -> torch.Tensor:
    """Counter-based N(0,1) noise on the device: fp32 [slices, B, n_per_utt] (include/stz.h: stz_philox_normal)."""
    lib = load_library()
    if out is None:
        out = torch.empty(slices, B, n_per_utt, dtype=torch.float32, device=torch.device("cuda", device))
    st = torch.cuda.current_stream(out.device).cuda_stream
    rc = lib.stz_philox_normal(int(seed) & (2 ** 64 - 1), int(first_utterance), slices, B, n_per_utt, _ptr(out),
                               out.device.index or 0, C.c_void_p(st))
    if rc != 0:
        raise StzError(f"stz_philox_normal failed ({rc}): {lib.stz_last_error(None).decode()}")
    return out


def sampler_plan(cfg: StzConfig, steps: int, sampler="student", cfg_scale: float = 1.0):
    """Host-only (runs without a GPU): the library's sigma schedule and fused sampler-step coefficient tables of one
    call (include/stz.h: stz_debug_plan) -> dict(sigma [E] f64, coef [E,8] f32, tfeat [E,d_time] f32, sigma0, cin0)."""
    lib = load_library()
    kind = _kind(sampler)
    E = 2 * steps if kind == SAMPLER_TEACHER else steps
    sigma = torch.empty(E, dtype=torch.float64)
    coef = torch.empty(E, 8, dtype=torch.float32)
    tfeat = torch.empty(E, cfg.d_time, dtype=torch.float32)
    init = torch.empty(2, dtype=torch.float64)
    rc = lib.stz_debug_plan(C.byref(_cconfig(cfg)), int(steps), kind, float(cfg_scale), _ptr(sigma), _ptr(coef), _ptr(tfeat),
                            _ptr(init))
    if rc != E:
        raise StzError(f"stz_debug_plan returned {rc}, expected {E} evaluations")
    return {"sigma": sigma, "coef": coef, "tfeat": tfeat, "sigma0": float(init[0]), "cin0": float(init[1])}
