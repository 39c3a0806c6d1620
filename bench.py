#!/usr/bin/env python
"""bench.py — the headline measurement of the StyleTTS-ZS hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], "cfg2"): per GPU, B = 64 utterances, T = 64 text tokens,
P = 50 prompt tokens, distilled student, 4 Euler steps with classifier-free guidance (cond/uncond
batched), K = 50 style codes x 512 channels, then the duration predictor on the sampled codes.
One "step" = one pass of that path over one batch.  Weak scaling: every rank runs its own batch,
no collective on the data path; value = utterances of all ranks / max-over-ranks time.

  value     inputs already resident in HBM, device-pointer C ABI (stz_sample_style +
            stz_predict_duration), CUDA events around each step, L2 flushed between steps.
  e2e       the same step through the host-buffer C ABI with pinned HOST buffers: H2D of the
            inputs, both kernels' work, D2H of style codes and durations, inside the timed region.
            Two pipeline slots (stz_synthesize_host_submit / _wait): batch i+1's H2D overlaps batch
            i's compute.  e2e_sync is the blocking one-call-per-step form (stz_synthesize_host).
  sustained the device-resident step in a loop for >= 2 s (power / thermal steady state).
  roofline  the tcgen05 GEMM family (dominant kernels) timed IN the real step: one CUDA-event pair per launch on the
            launching stream (library "profile" mode, eager), algorithmic 2MNK flops / summed time vs the measured bf16
            peak.  Sub-fields: the whole evaluation loop (graph) against its algorithmic flops, the isolated-kernel
            microbenchmark, and the committed ncu evidence (profiles/): launch-list share and tensor-pipe %.
  configs   (N = 1) BASELINE configs 1, 3, 4 at their full sizes, each with its roofline fraction and the oracle timed on
            a stated sub-sample on the same box.
  sharded   strong scaling through the sharder (shard.py): ONE global variable-length batch of 1024 utterances split over
            the ranks, per-rank host-buffer call, results gathered on the host (shared mapping), all inside the timed region.
  cpu_baseline  oracle/ (fp32 PyTorch restatement) on the host cores, same workload, one batch.

`--impl reference` times the reference arm.  The reference ships no implementation of this path
(/root/reference/README.md:15-16), so that arm is the oracle port on the host CPU cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "style-sampled utterances/sec (distilled 4-step CFG sampler + duration predictor)"
UNIT = "utterances/s"
WORK = dict(B=64, T=64, P=50, steps=4, cfg_scale=2.0, sampler="student")
FRAMES_PER_S = 80.0  # 24 kHz / hop 300
SHARDED = dict(B=1024, T=512, lens=(16, 512), steps=4)     # the strong-scaling leg's global batch (BASELINE configs[3]/[4] shape)


def workload_config(n_gpus: int):
    return {"workload": "cfg2: batch 64 per GPU, 64 text tokens, 50 prompt tokens, distilled 4-step Euler sampler with "
                        "CFG (cond/uncond batched, 8 denoiser sequence-evals per utterance), 50x512 style codes, "
                        "+ duration predictor (4 BiLSTM layers)",
            "per_gpu_batch": WORK["B"], "global_batch": WORK["B"] * n_gpus, "text_tokens": WORK["T"],
            "sampler_steps": WORK["steps"], "cfg_scale": WORK["cfg_scale"],
            "parallelism": f"utterance-sharded x{n_gpus}, no collective",
            "l2": "flushed between timed steps (256 MiB memset outside the event pairs)"}


# ----------------------------------------------------------------------------------------------
# algorithmic work
# ----------------------------------------------------------------------------------------------
def sampler_flops(cfg, B: int, T, P: int, evals: int, branches: int = 2) -> float:
    """Algorithmic flops of one sample_style call.  T: an int or a list of per-utterance valid lengths (padding is not
    credited).  Per sequence-evaluation: L [2K(6d^2 + 2 d d_ff) + 4Kd(K + S) + 18 d^2] + 4 K Ds d + 4 d^2, S = T + P for
    the conditional branch, T + 1 for the unconditional one; once per utterance: 2 S d^2 + L 4 S d^2 (context tokens)."""
    d, L, K, Ds, dff = cfg.d_model, cfg.n_layers, cfg.n_style, cfg.d_style, cfg.d_ff
    lens = [T] * B if isinstance(T, int) else list(T)
    total = 0.0
    for t in lens:
        for S in (t + P, t + 1)[:branches]:      # branches = 1: the guidance-conditioned student (conditional branch only)
            total += evals * (L * (2 * K * (6 * d * d + 2 * d * dff) + 4 * K * d * (K + S) + 18 * d * d) + 4 * K * Ds * d + 4 * d * d)
        S = t + P
        total += 2 * S * d * d + L * 4 * S * d * d
    return total


def load_json(*parts):
    try:
        with open(os.path.join(ROOT, *parts)) as f:
            return json.load(f)
    except Exception:
        return None


def peaks():
    pk = load_json("MEASURED_PEAKS.json") or {}
    if "bf16_tflops" in pk:
        return {"burst": pk["bf16_tflops"], "sustained": pk.get("bf16_tflops_sustained", pk["bf16_tflops"]),
                "hbm_gbs": pk.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        return len(self.rows)

    def summary(self, lo: int = 0, hi=None):
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows[lo:hi]:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        return self.summary()


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on the host cores
# ----------------------------------------------------------------------------------------------
def oracle_step(oracle, inp, steps, sampler, with_predictor=True):
    z = oracle.sample_style(inp["text_emb"], inp["prompt_feats"], steps, WORK["cfg_scale"],
                            text_mask=inp["text_mask"], noise=inp["noise"], sampler=sampler)
    d = oracle.predict_duration(inp["text_emb"], z, text_mask=inp["text_mask"]) if with_predictor else None
    return z, d


def make_oracle():
    import torch
    import styletts_zs_b200 as stz
    from oracle.model import OraclePath  # the one sanctioned non-test use: cpu_baseline / --impl reference
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = stz.DEFAULT
    return OraclePath(cfg, stz.init_weights(cfg, 0)), torch.get_num_threads()


def run_oracle(batch: int, steps: int, warmup: int):
    import styletts_zs_b200 as stz
    o, threads = make_oracle()
    cfg = stz.DEFAULT
    inp = stz.synthetic_inputs(cfg, batch, WORK["T"], steps=WORK["steps"], seed=1234)
    for _ in range(warmup):
        oracle_step(o, stz.synthetic_inputs(cfg, min(batch, 2), WORK["T"], steps=WORK["steps"], seed=1), WORK["steps"], WORK["sampler"])
    t0 = time.perf_counter()
    frames = 0
    for _ in range(steps):
        _, d = oracle_step(o, inp, WORK["steps"], WORK["sampler"])
        frames += int(d.sum())
    dt = time.perf_counter() - t0
    return dict(utt_per_s=batch * steps / dt, ms_per_step=dt / steps * 1e3, cores=os.cpu_count() or 1, frames=frames, wall_s=dt,
                threads=threads)


def main_reference(args, rank):
    if rank != 0:
        return 0
    batch = WORK["B"]      # the full cfg2 batch: ~1.2 s per step on 16 host cores, the CPU's most efficient operating point
    r = run_oracle(batch, args.steps, max(args.warmup, 1))
    sample = (f"{batch} utterances per step = one full cfg2 batch (T=64, 4-step CFG student + predictor), fp32 PyTorch "
              f"oracle port, {r['threads']} threads")
    line = {"impl": "reference", "metric": METRIC, "value": r["utt_per_s"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": r["utt_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": sample},
            "e2e": {"value": r["utt_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference ships no code (README.md:15-16 'under construction'); this arm is the oracle port on host cores"}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
# native arm: legs
# ----------------------------------------------------------------------------------------------
def timed_steps(torch, fn, n, flush):
    """n calls of fn, one CUDA-event pair each, L2 flushed between them -> list of ms."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in ev]


def roofline_leg(torch, stz, path, cfg, dev_step, sample_only, B, T, steps, flush, n_timed):
    """Roofline of the dominant kernel family (tcgen05 GEMMs of the denoiser)."""
    pk = peaks()
    # (1) in the real step: eager profile mode, one event pair per launch on the launching stream
    path.set_option("profile", 1)
    reps = 3
    for _ in range(reps):
        dev_step()
    prof = path.profile_read()
    path.set_option("profile", 0)
    ms, flops, n = prof["gemm_tc"]
    ach = flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
    tot_ms = sum(v[0] for v in prof.values())
    # (2) the whole evaluation loop (one CUDA graph + conditioning prep) against its algorithmic flops
    s_ms = timed_steps(torch, sample_only, n_timed, flush)
    s_ms = sum(s_ms) / len(s_ms)
    s_fl = sampler_flops(cfg, B, T, cfg.n_style, steps)
    whole = {"ms": s_ms, "algorithmic_tflop": s_fl / 1e12, "achieved": s_fl / (s_ms * 1e-3) / 1e12}
    whole["frac_of_burst"] = whole["achieved"] / pk["burst"]
    whole["frac_of_sustained"] = whole["achieved"] / pk["sustained"]
    whole["note"] = ("sample_style as the caller sees it (graph replay, PDL intact), ALL kernels' time (attention, prep, "
                     "LayerNorm passes included) against the algorithmic flops of SURVEY.md §8d: a lower bound of the GEMM "
                     "family's in-graph efficiency")
    # (3) isolated microbenchmark per GEMM shape (random operands, L2-warm, back-to-back)
    R, L, d, dff = 2 * B * cfg.n_style, cfg.n_layers, cfg.d_model, cfg.d_ff
    fused = path.get_option("last_fuse_mode") == 3
    res_epi = 6 if fused else 4
    mix = [("qkv", 3 * d, d, 2, L), ("attn_out" + ("+ln" if fused else ""), d, d, res_epi, L), ("q_cross", d, d, 2, L),
           ("cross_out" + ("+ln" if fused else ""), d, d, res_epi, L), ("ffn1_gelu", dff, d, 3, L),
           ("ffn2" + ("+ln" if fused else ""), d, dff, res_epi, L)]
    per_shape, tot_flops, tot_us = {}, 0.0, 0.0
    for name, N, K, epi, cnt in mix:
        us = path.bench_gemm(R, N, K, epi, 50)
        fl = 2.0 * R * N * K
        per_shape[name] = {"M": R, "N": N, "K": K, "us": round(us, 2), "tflops": round(fl / us * 1e-6, 1)}
        tot_flops += fl * cnt
        tot_us += us * cnt
    iso = tot_flops / tot_us * 1e-6
    ncu = load_json("profiles", "r02_roofline_evidence.json")
    return {"kernel": "gemm2_kernel / gemmln3_kernel (persistent tcgen05/TMEM/TMA bf16 GEMM family of the denoiser)",
            "bound": "tensor", "achieved": ach, "peak": pk["burst"], "unit": "TFLOP/s", "frac": ach / pk["burst"],
            "frac_of_sustained": ach / pk["sustained"], "peak_sustained": pk["sustained"], "peak_source": pk["source"],
            "traffic": (ncu or {}).get("gemm_dram_bytes_per_launch"),
            "avg_launch_us": ms / max(n, 1) * 1e3, "launches_per_step": n // reps,
            "flops_per_launch": flops / max(n, 1),
            "how": "IN the real step: library profile mode runs the step eagerly with one CUDA-event pair around every launch "
                   "on the launching stream; achieved = sum(2MNK of the tcgen05 GEMM launches) / sum(their event times), "
                   f"{reps} steps.  Event pairs add launch gaps and lose the PDL overlap the graph has, so this is the "
                   "conservative reading; `whole_sampler` is the graph as timed, `isolated_microbench` the optimistic one.",
            "share_of_profiled_step": ms / tot_ms if tot_ms > 0 else None,
            "classes_ms_per_profiled_step": {k: v[0] / reps for k, v in prof.items()},
            "whole_sampler": whole,
            "isolated_microbench": {"achieved": iso, "frac_of_burst": iso / pk["burst"], "per_shape": per_shape,
                                    "note": "50 back-to-back launches per shape, random operands, L2-warm, no data dependency "
                                            "between launches: an upper bound, not the in-step number"},
            "ncu_evidence": ncu}


def predictor_leg(torch, path, cfg, dev, z, B, T, flush, n_timed):
    """The duration predictor on its own (report achieved GB/s AND fp32 FLOP/s, say which binds)."""
    pred = lambda: path.predict_duration(dev["text_emb"], z)
    pred_ms = timed_steps(torch, pred, n_timed, flush)
    pred_ms = sum(pred_ms) / len(pred_ms)
    path.set_option("profile", 1)
    reps = 3
    for _ in range(reps):
        pred()
    prof = path.profile_read()
    path.set_option("profile", 0)
    h, dh, ds, nl = cfg.h_lstm, cfg.d_hid, cfg.d_sty_tok, cfg.n_lstm
    fl_tok = nl * 2 * (2 * 4 * h * (dh + ds) + 2 * 4 * h * h) + (nl - 1) * 2 * ds * 2 * dh \
        + 2 * cfg.d_text * ds + 4 * cfg.n_style * ds + 2 * ds * ds + 2 * dh * cfg.max_dur
    ntok = B * T
    ew_ms, ew_bytes, ew_n = prof["pred_ew"]
    cls = {}
    for k, (ms, work, n) in prof.items():
        if n:
            cls[k] = {"ms": ms / reps, "launches": n // reps, "work": work / reps,
                      "rate": (work / (ms * 1e-3) / (1e9 if k in ("pred_ew", "ln_mod") else 1e12)) if ms > 0 else None,
                      "rate_unit": "GB/s" if k in ("pred_ew", "ln_mod") else "TFLOP/s"}
    pk = peaks()
    ncu = load_json("profiles", "r02_ncu_predictor.json")
    return {"ms": pred_ms, "tokens": ntok, "fp32_flops_per_token": fl_tok,
            "achieved_tflops_fp32_equiv": ntok * fl_tok / (pred_ms * 1e-3) / 1e12,
            "elementwise_bytes_per_token": ew_bytes / reps / ntok if ntok else None,
            "achieved_gbs": ew_bytes / (ew_ms * 1e-3) / 1e9 if ew_ms > 0 else None,
            "achieved_gbs_frac_of_hbm_peak": (ew_bytes / (ew_ms * 1e-3) / 1e9 / pk["hbm_gbs"]) if ew_ms > 0 and pk["hbm_gbs"] else None,
            "how": "predict_duration alone, CUDA events; per class: library profile mode (one event pair per launch); "
                   "achieved_gbs = algorithmic bytes of the fp32 elementwise kernels (each input + output tensor once) / "
                   "their summed event time; measured dram__bytes per kernel: ncu_dram (profiles/r02_ncu_predictor.json)",
            "classes": cls, "ncu_dram": ncu,
            "binds": "neither HBM nor FMA throughput: the BiLSTM's serial chain (T steps x n_lstm layers x ~1.15 us per "
                     "recurrent step: DSMEM exchange of h_t across the 8-CTA cluster ~1.1 k cycles + 32 tcgen05.mma with W_hh "
                     "in tensor memory ~0.5 k + gate math; profiles/r02_ab_lstm_forms.txt); "
                     "the working set (< 40 MB at cfg2) is L2-resident, so the elementwise kernels' GB/s is L2 traffic",
            "chain_us": T * nl * 1.15}


def config_legs(torch, stz, path, cfg, flush):
    """BASELINE configs 1, 3, 4 at full size on this GPU, with the oracle timed on a stated sub-sample (N = 1 only)."""
    pk = peaks()
    oracle, threads = make_oracle()
    out = {}
    specs = {
        "cfg1": dict(B=1, T=64, steps=1, sampler="student", var_len=None, iters=8, oracle_B=1, pred=True,
                     desc="1 utterance, distilled 1-step sampling + duration predictor"),
        "cfg3": dict(B=32, T=64, steps=32, sampler="teacher", var_len=None, iters=4, oracle_B=2, pred=False,
                     desc="undistilled teacher: 32 ADPM2 steps (64 CFG evaluations), batch 32"),
        "cfg4": dict(B=256, T=512, steps=4, sampler="student", var_len=(16, 512), iters=4, oracle_B=16, pred=True,
                     desc="variable-length text (16..512 tokens, padding masks), batch 256, distilled 4-step sampler + predictor"),
        # the cfg2 workload with the guidance-conditioned student (one branch per step)
        "cfg2_guided_student": dict(B=64, T=64, steps=4, sampler="guided", var_len=None, iters=8, oracle_B=16, pred=True,
                                    desc="cfg2's batch with the guidance-conditioned student: guidance scale as an input "
                                         "embedding, ONE branch per step (4 denoiser sequence-evals per utterance) + predictor"),
    }
    for name, sp in specs.items():
        kind = {"teacher": stz.SAMPLER_TEACHER, "guided": stz.SAMPLER_GUIDED}.get(sp["sampler"], stz.SAMPLER_STUDENT)
        inp = stz.synthetic_inputs(cfg, sp["B"], sp["T"], steps=sp["steps"], sampler=kind, seed=1234, var_len=sp["var_len"])
        dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}
        tm = inp["text_mask"].cuda() if sp["var_len"] else None
        last = {}

        def samp():
            last["z"] = path.sample_style(dev["text_emb"], dev["prompt_feats"], sp["steps"], WORK["cfg_scale"], noise=dev["noise"],
                                          text_mask=tm, sampler=sp["sampler"])

        def step():
            samp()
            if sp["pred"]:
                last["d"] = path.predict_duration(dev["text_emb"], last["z"], text_mask=tm)
        for _ in range(2):
            step()
        ms = timed_steps(torch, step, sp["iters"], flush)
        s_ms = timed_steps(torch, samp, sp["iters"], flush)
        ms, s_ms = sum(ms) / len(ms), sum(s_ms) / len(s_ms)
        E = 2 * sp["steps"] if sp["sampler"] == "teacher" else sp["steps"]
        fl = sampler_flops(cfg, sp["B"], inp["lens"].tolist() if sp["var_len"] else sp["T"], cfg.n_style, E,
                           branches=1 if sp["sampler"] == "guided" else 2)
        frames = int(last["d"].sum()) if sp["pred"] else None
        # the oracle on the first oracle_B utterances of the same batch (utterances are independent)
        nb = sp["oracle_B"]
        tmax = int(inp["lens"][:nb].max())
        sub = {"text_emb": inp["text_emb"][:nb, :tmax], "prompt_feats": inp["prompt_feats"][:nb], "noise": inp["noise"][:, :nb],
               "text_mask": inp["text_mask"][:nb, :tmax]}
        t0 = time.perf_counter()
        oracle_step(oracle, sub, sp["steps"], sp["sampler"], with_predictor=sp["pred"])
        o_s = time.perf_counter() - t0
        ach = fl / (s_ms * 1e-3) / 1e12
        out[name] = {"workload": sp["desc"], "B": sp["B"], "T": sp["T"], "ms": ms, "utt_per_s": sp["B"] / ms * 1e3,
                     "sampler_ms": s_ms, "path_rtf": (ms * 1e-3) / (frames / FRAMES_PER_S) if frames else None,
                     "fused_gemm_adaln": path.get_option("last_fuse_mode") == 3,
                     "roofline": {"bound": "tensor", "algorithmic_tflop": fl / 1e12, "achieved": ach, "unit": "TFLOP/s",
                                  "frac_of_burst": ach / pk["burst"], "frac_of_sustained": ach / pk["sustained"],
                                  "floor_ms_at_sustained_peak": fl / (pk["sustained"] * 1e12) * 1e3,
                                  "note": "the whole sample_style call (graph + prep) against its algorithmic flops"},
                     "cpu_oracle": {"utt_per_s": nb / o_s, "seconds": o_s, "threads": threads,
                                    "sample": f"the first {nb} of the batch's {sp['B']} utterances, one pass, fp32 PyTorch oracle"},
                     "speedup_vs_cpu_oracle": (sp["B"] / ms * 1e3) / (nb / o_s)}
    # the prosody heads (F0 / energy over the length-regulated frames) on cfg2's batch
    inp = stz.synthetic_inputs(cfg, WORK["B"], WORK["T"], steps=1, seed=1234)
    text = inp["text_emb"].cuda()
    style = (0.7 * torch.randn(WORK["B"], cfg.n_style, cfg.d_style, generator=torch.Generator().manual_seed(5))).cuda()
    F_max = 1536
    pros = lambda: last.__setitem__("p", path.predict_prosody(text, style, max_frames=F_max))
    last = {}
    for _ in range(2):
        pros()
    ms = timed_steps(torch, pros, 5, flush)
    ms = sum(ms) / len(ms)
    frames = int(last["p"][2].sum())
    nb = 4
    t0 = time.perf_counter()
    oracle.predict_prosody(inp["text_emb"][:nb], style[:nb].cpu(), max_frames=F_max)
    o_s = time.perf_counter() - t0
    out["prosody_heads"] = {"workload": f"predict_prosody on cfg2's batch: duration predictor + length regulator + BiLSTM over the frames + F0 / "
                                        f"energy heads, {WORK['B']} utterances, {frames} frames in total (F_max {F_max})",
                            "ms": ms, "utt_per_s": WORK["B"] / ms * 1e3, "frames_per_s": frames / ms * 1e3,
                            "path_rtf": (ms * 1e-3) / (frames / FRAMES_PER_S) if frames else None,
                            "cpu_oracle": {"utt_per_s": nb / o_s, "seconds": o_s, "threads": threads,
                                           "sample": f"the first {nb} utterances, one pass, fp32 PyTorch oracle"},
                            "bound": "the frame BiLSTM's serial chain: longest utterance's frame count x ~1.2 us per recurrent step"}
    return out


def sharded_leg(torch, stz, path, cfg, rank, world, local_rank, barrier, reps=3):
    """Strong scaling through the sharder: one global variable-length batch, length-sorted round-robin over the ranks, the
    per-rank host-buffer call, results gathered in a shared host mapping, every rank re-ordering its own rows — all timed."""
    Bg, Tg, steps = SHARDED["B"], SHARDED["T"], SHARDED["steps"]
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))     # torchrun pins OMP_NUM_THREADS=1: the re-ordering copy is host work
    g = torch.Generator().manual_seed(4321)
    lens = torch.randint(SHARDED["lens"][0], SHARDED["lens"][1] + 1, (Bg,), generator=g)
    shards = stz.shard_utterances(lens.tolist(), world)
    shard_T = [max(int(lens[i]) for i in sh) if sh else 1 for sh in shards]
    idx = shards[rank]
    n, t = len(idx), shard_T[rank]
    gr = torch.Generator().manual_seed(99 + rank)           # values differ per rank; the WORK (lengths) is the global batch's
    my_lens = lens[torch.tensor(idx)]
    text = torch.randn(n, t, cfg.d_text, generator=gr).pin_memory()
    prompt = torch.randn(n, cfg.n_style, cfg.d_prompt, generator=gr).pin_memory()
    mask = (torch.arange(t)[None] < my_lens[:, None])
    tag = os.environ.get("MASTER_PORT", "0") + f"_{os.getppid() if world > 1 else os.getpid()}"
    out = stz.SharedHostOutputs(tag, Bg, Tg, cfg.n_style, cfg.d_style, rank, world, barrier)
    sh_in = {"text_emb": text, "text_mask": mask, "prompt_feats": prompt}

    def compute(te, tm, pf, pm, nz, out_style=None, out_dur=None):
        return path.synthesize_host(te, pf, steps, WORK["cfg_scale"], text_mask=tm, seed=2024, first_utterance=list(idx),
                                    out_style=out_style, out_dur=out_dur)
    res = None
    for _ in range(2):                                       # warm-up: workspace growth, graph capture at this shape
        res = stz.synthesize_sharded_shm(compute, sh_in, shards, shard_T, out, barrier)
    times, t_compute, t_asm = [], [], []
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        o_style, o_dur = out.slab(n, t)
        compute(text, mask, prompt, None, None, out_style=o_style, out_dur=o_dur)
        t1 = time.perf_counter()
        out.scatter_own(shards, shard_T)
        t2 = time.perf_counter()
        barrier()
        if rank == 0:
            res = out.ordered()
        t3 = time.perf_counter()
        times.append(t3 - t0); t_compute.append(t1 - t0); t_asm.append(t2 - t1)
    ok = None
    if rank == 0:
        style, dur = res
        valid = torch.arange(Tg)[None] < lens[:, None]
        ok = bool(torch.isfinite(style).all()) and bool((dur[valid] >= 1).all()) and bool((dur[~valid] == 0).all())
    tt = torch.tensor([min(times), min(t_compute), min(t_asm)], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total, comp, asm = (float(x) for x in tt)
    out.close(barrier)
    return {"workload": f"ONE global batch of {Bg} utterances, text lengths uniform in [{SHARDED['lens'][0]}, {SHARDED['lens'][1]}] (padding "
                        f"masks), {steps}-step CFG student + duration predictor, length-sorted round-robin over {world} rank(s)",
            "scaling": "strong", "global_batch": Bg, "per_rank_batch": n, "value": Bg / total, "unit": UNIT, "seconds": total,
            "per_rank_call_seconds_max": comp, "per_rank_reorder_seconds_max": asm,
            "host_gather": "every rank's D2H lands in its slab of one /dev/shm mapping registered as pinned memory "
                           f"(pinned={out.pinned}); every rank copies ITS rows to their places in the caller's utterance order "
                           f"inside the same mapping ({Bg * cfg.n_style * cfg.d_style * 4 / 1e6:.0f} MB of style codes + durations "
                           "in total, 1 / world of it per rank, in parallel); barrier; rank 0 returns views: no collective, no "
                           "pickling, no serial re-ordering pass",
            "h2d_bytes_per_rank": int(text.numel() * 4 + prompt.numel() * 4 + mask.numel()),
            "d2h_bytes_per_rank": int(n * cfg.n_style * cfg.d_style * 4 + n * t * 4), "results_ok": ok,
            "timing": "wall clock around [per-rank blocking stz_synthesize_host into the shared slab; per-rank re-order; barrier], "
                      f"best of {reps}, max over ranks; inputs are each rank's own pinned host tensors (a server hands every "
                      "rank its utterances), noise drawn on the device from global utterance indices"}


def main_native(args, rank, world, local_rank):
    import torch
    import styletts_zs_b200 as stz

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        # NCCL prints its version banner on stdout at communicator creation: keep stdout to the one JSON line by pointing
        # fd 1 at stderr until the communicator exists
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = stz.DEFAULT
    B, T, steps, scale = WORK["B"], WORK["T"], WORK["steps"], WORK["cfg_scale"]
    path = stz.StyleTTSZSPath(cfg, stz.init_weights(cfg, 0), device=local_rank)
    inp = stz.synthetic_inputs(cfg, B, T, steps=steps, seed=1234 + rank)
    dev = {k: inp[k].cuda() for k in ("text_emb", "prompt_feats", "noise")}
    host = {k: inp[k].pin_memory() for k in ("text_emb", "prompt_feats", "noise")}
    out_style = torch.empty(B, cfg.n_style, cfg.d_style).pin_memory()
    out_dur = torch.empty(B, T, dtype=torch.int32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    last = {}

    def sample_only():
        last["z"] = path.sample_style(dev["text_emb"], dev["prompt_feats"], steps, scale, noise=dev["noise"])
        return last["z"]

    def dev_step():
        z = sample_only()
        last["d"] = path.predict_duration(dev["text_emb"], z)
        return z, last["d"]

    def host_step():
        return path.synthesize_host(host["text_emb"], host["prompt_feats"], steps, scale, noise=host["noise"],
                                    out_style=out_style, out_dur=out_dur)

    # two pipeline slots, each with its own pinned host buffers (stz_synthesize_host_submit / _wait)
    host2 = [host, {k: inp[k].clone().pin_memory() for k in ("text_emb", "prompt_feats", "noise")}]
    outs2 = [(out_style, out_dur), (torch.empty_like(out_style).pin_memory(), torch.empty_like(out_dur).pin_memory())]

    def submit(slot):
        h = host2[slot]
        path.synthesize_host(h["text_emb"], h["prompt_feats"], steps, scale, noise=h["noise"], out_style=outs2[slot][0],
                             out_dur=outs2[slot][1], slot=slot)

    def host_step_seeded():   # same call, noise drawn on the device (Philox, bit-identical to oracle/philox.py): no noise H2D
        return path.synthesize_host(host["text_emb"], host["prompt_feats"], steps, scale, seed=1234, first_utterance=rank * B,
                                    out_style=out_style, out_dur=out_dur)

    # ---- warm-up (graph capture, workspace growth) --------------------------------------------
    for _ in range(max(args.warmup, 3)):
        dev_step()
        host_step()
        host_step_seeded()
        submit(0); submit(1); path.synthesize_host_wait(0); path.synthesize_host_wait(1)
    barrier()

    # ---- value: device-resident inputs ---------------------------------------------------------
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    n0 = path.launch_count()
    barrier()
    for i, (a, b) in enumerate(ev):
        flush.zero_()
        if args.ncu and i == len(ev) - 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()      # ncu --profile-from-start off: capture exactly one timed step
        a.record()
        z, d = dev_step()
        b.record()
        if args.ncu and i == len(ev) - 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
    barrier()
    launches = path.launch_count() - n0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    frames = int(d.sum())

    # ---- e2e: host buffers through the C ABI ----------------------------------------------------
    barrier()
    e2e_s = 0.0
    for _ in range(1 if args.ncu else args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        host_step()
        e2e_s += time.perf_counter() - t0
    barrier()
    # pipelined e2e: batch i+1 is submitted (its H2D copies start) before batch i is waited for; every batch's H2D and D2H
    # are inside the timed region; the L2 flush of each step runs (untimed-for, but inside the region) on torch's stream
    e2e_pipe_s = 0.0
    if not args.ncu:
        barrier()
        t0 = time.perf_counter()
        flush.zero_()
        submit(0)
        for i in range(args.steps):
            if i + 1 < args.steps:
                flush.zero_()
                submit((i + 1) & 1)
            path.synthesize_host_wait(i & 1)
        torch.cuda.synchronize()
        e2e_pipe_s = time.perf_counter() - t0
        barrier()
    e2e_seed_s = 0.0
    for _ in range(1 if args.ncu else args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        host_step_seeded()
        e2e_seed_s += time.perf_counter() - t0
    barrier()
    clk_mark = clocks.mark() if rank == 0 else 0
    clk = clocks.summary(0, clk_mark) if rank == 0 else None

    # ---- sustained: >= 2 s of back-to-back device-resident steps (no flush: the ~230 MB working set exceeds the 126 MB L2)
    sus_ms, sus_n = 0.0, 0
    if not args.ncu:
        sus_n = max(50, int(args.sustain_s * 1e3 / max(dev_ms / args.steps, 1e-3)))
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(sus_n):
            dev_step()
        b.record()
        barrier()
        sus_ms = a.elapsed_time(b)
    clk_sus = clocks.summary(clk_mark, None) if rank == 0 else None

    t = torch.tensor([dev_ms, e2e_s * 1e3, e2e_seed_s * 1e3, e2e_pipe_s * 1e3, sus_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_sync_ms, e2e_seed_ms, e2e_pipe_ms, sus_ms = (float(x) for x in t)
    e2e_ms = e2e_pipe_ms if e2e_pipe_ms > 0 else e2e_sync_ms
    total_utt = B * world * args.steps
    h2d = sum(host[k].numel() * 4 for k in host)
    d2h = out_style.numel() * 4 + out_dur.numel() * 4

    roofline, cpu, predictor, configs, sharded = None, None, None, None, None
    if rank == 0 and not args.ncu:
        n_timed = max(5, min(args.steps, 20))
        predictor = predictor_leg(torch, path, cfg, dev, last["z"], B, T, flush, n_timed)
        roofline = roofline_leg(torch, stz, path, cfg, dev_step, sample_only, B, T, steps, flush, n_timed)
        if world == 1:
            r = run_oracle(B, 1, 1)
            cpu = {"value": r["utt_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                   "sample": f"one full cfg2 batch ({B} utterances) through the fp32 PyTorch oracle, {r['wall_s']:.1f} s wall"}
            if not args.no_configs:
                try:
                    configs = config_legs(torch, stz, path, cfg, flush)
                except Exception as e:      # a supplementary leg must not cost the headline line
                    configs = {"error": f"{type(e).__name__}: {e}"}
    if not args.ncu and not args.no_sharded:
        try:
            sharded = sharded_leg(torch, stz, path, cfg, rank, world, local_rank, barrier)
        except Exception as e:
            sharded = {"error": f"{type(e).__name__}: {e}"}
    clk_all = clocks.stop() if rank == 0 else None
    line = {"metric": METRIC, "value": total_utt / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "dtype_note": "bf16 tensor-core operands (split-bf16 = fp32-grade where noted in DESIGN.md), fp32 accumulate / residual / "
                          "sampler state / LSTM gates / duration head",
            "data": "synthetic (seeded N(0,1) inputs, random-init weights)", "config": workload_config(world),
            "e2e": {"value": total_utt / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "pipeline_depth": 2 if e2e_pipe_ms > 0 else 1,
                    "how": "one blocking stz_synthesize_host call per step" if e2e_pipe_ms <= 0 else
                           "stz_synthesize_host_submit / _wait on two slots with pinned HOST buffers: batch i+1 is submitted "
                           "before batch i is waited for, so its H2D overlaps batch i's compute; every step's H2D, compute, "
                           "D2H and the per-step L2 flush are inside the timed region (wall clock, max over ranks)"},
            "e2e_sync": {"value": total_utt / (e2e_sync_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                         "ms_per_step": e2e_sync_ms / args.steps,
                         "note": "one blocking stz_synthesize_host call per step (no overlap between steps)"},
            "e2e_device_noise": {"value": total_utt / (e2e_seed_ms * 1e-3), "unit": UNIT,
                                 "h2d_bytes_per_step": h2d - host["noise"].numel() * 4, "d2h_bytes_per_step": d2h,
                                 "ms_per_step": e2e_seed_ms / args.steps,
                                 "note": "the blocking e2e call with seed= instead of a host noise tensor: noise drawn on the device "
                                         "(Philox4x32-10, bit-identical to oracle/philox.py); informational, `e2e` is the headline"},
            "sustained": None if sus_n == 0 else
                         {"value": B * world * sus_n / (sus_ms * 1e-3), "unit": UNIT, "seconds": sus_ms * 1e-3, "steps": sus_n,
                          "ms_per_step": sus_ms / sus_n, "clocks": clk_sus,
                          "note": "device-resident steps back to back between one CUDA-event pair (max over ranks), no L2 flush "
                                  "(the step's working set exceeds L2): the steady-state rate beside the short timed region"},
            "gpu_launches": int(launches), "clocks": clk, "clocks_whole_run": clk_all, "roofline": roofline, "cpu_baseline": cpu,
            "predictor": predictor, "configs": configs, "sharded": sharded,
            "path_rtf": (dev_ms * 1e-3 / args.steps) / (frames / FRAMES_PER_S) if frames else None}
    if rank == 0:
        print(json.dumps(line), flush=True)
    path.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--sustain-s", type=float, default=2.5, dest="sustain_s", help="length of the sustained leg (seconds)")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg1 / cfg3 / cfg4 legs (N = 1)")
    ap.add_argument("--no-sharded", action="store_true", help="skip the strong-scaling sharder leg")
    ap.add_argument("--ncu", action="store_true",
                    help="profiling run: cudaProfilerStart/Stop around the last timed step, skip every other leg "
                         "(numbers printed by such a run are not bench values)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return main_reference(args, rank)
    return main_native(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
