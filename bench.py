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

  value  : inputs already resident in HBM, device-pointer C ABI (stz_sample_style +
           stz_predict_duration), CUDA events around each step, L2 flushed between steps.
  e2e    : the same step through the host-buffer C ABI with pinned HOST buffers: H2D of the
           inputs, both kernels' work, D2H of style codes and durations, inside the timed region.
           Two pipeline slots (stz_synthesize_host_submit / _wait): batch i+1's H2D overlaps batch
           i's compute.  e2e_sync is the blocking one-call-per-step form (stz_synthesize_host).
  roofline: the tcgen05 GEMM family (dominant kernel) timed in situ with a CUDA-event pair per
           launch (library "profile" mode), algorithmic flops / summed time vs measured bf16 peak.
  cpu_baseline: oracle/ (fp32 PyTorch restatement) on the host cores, same workload, one batch.

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


def workload_config(n_gpus: int):
    return {"workload": "cfg2: batch 64 per GPU, 64 text tokens, 50 prompt tokens, distilled 4-step Euler sampler with "
                        "CFG (cond/uncond batched, 8 denoiser sequence-evals per utterance), 50x512 style codes, "
                        "+ duration predictor (4 BiLSTM layers)",
            "per_gpu_batch": WORK["B"], "global_batch": WORK["B"] * n_gpus, "text_tokens": WORK["T"],
            "sampler_steps": WORK["steps"], "cfg_scale": WORK["cfg_scale"],
            "parallelism": f"utterance-sharded x{n_gpus}, no collective",
            "l2": "flushed between timed steps (256 MiB memset outside the event pairs)"}


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

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on the host cores
# ----------------------------------------------------------------------------------------------
def oracle_step(oracle, inp):
    z = oracle.sample_style(inp["text_emb"], inp["prompt_feats"], WORK["steps"], WORK["cfg_scale"],
                            text_mask=inp["text_mask"], noise=inp["noise"], sampler=WORK["sampler"])
    d = oracle.predict_duration(inp["text_emb"], z, text_mask=inp["text_mask"])
    return z, d


def run_oracle(batch: int, steps: int, warmup: int):
    import torch
    import styletts_zs_b200 as stz
    from oracle.model import OraclePath  # the one sanctioned non-test use: cpu_baseline / --impl reference
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = stz.DEFAULT
    o = OraclePath(cfg, stz.init_weights(cfg, 0))
    inp = stz.synthetic_inputs(cfg, batch, WORK["T"], steps=WORK["steps"], seed=1234)
    for _ in range(warmup):
        oracle_step(o, stz.synthetic_inputs(cfg, min(batch, 2), WORK["T"], steps=WORK["steps"], seed=1))
    t0 = time.perf_counter()
    frames = 0
    for _ in range(steps):
        _, d = oracle_step(o, inp)
        frames += int(d.sum())
    dt = time.perf_counter() - t0
    return dict(utt_per_s=batch * steps / dt, ms_per_step=dt / steps * 1e3, cores=cores, frames=frames, wall_s=dt,
                threads=torch.get_num_threads())


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
# native arm
# ----------------------------------------------------------------------------------------------
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
        os.environ.setdefault("NCCL_DEBUG", "WARN")     # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

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

    def dev_step():
        z = path.sample_style(dev["text_emb"], dev["prompt_feats"], steps, scale, noise=dev["noise"])
        return z, path.predict_duration(dev["text_emb"], z)

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
    clk = clocks.stop() if rank == 0 else None
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

    t = torch.tensor([dev_ms, e2e_s * 1e3, e2e_seed_s * 1e3, e2e_pipe_s * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_sync_ms, e2e_seed_ms, e2e_pipe_ms = (float(x) for x in t)
    e2e_ms = e2e_pipe_ms if e2e_pipe_ms > 0 else e2e_sync_ms
    total_utt = B * world * args.steps
    h2d = sum(host[k].numel() * 4 for k in host)
    d2h = out_style.numel() * 4 + out_dur.numel() * 4

    roofline, cpu, predictor = None, None, None
    if rank == 0 and not args.ncu:
        # ---- the duration predictor on its own (report achieved GB/s AND fp32 FLOP/s, say which binds) ----
        pe = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in pe:
            flush.zero_()
            a.record()
            path.predict_duration(dev["text_emb"], z)
            b.record()
        torch.cuda.synchronize()
        pred_ms = sum(a.elapsed_time(b) for a, b in pe) / len(pe)
        h, dh, ds, nl = cfg.h_lstm, cfg.d_hid, cfg.d_sty_tok, cfg.n_lstm
        fl_tok = nl * 2 * (2 * 4 * h * (dh + ds) + 2 * 4 * h * h) + (nl - 1) * 2 * ds * 2 * dh \
            + 2 * cfg.d_text * ds + 4 * cfg.n_style * ds + 2 * ds * ds + 2 * dh * cfg.max_dur
        by_tok = 100e3   # <~100 KB of fp32 activations per token (each [rows, C] activation read + written once per pass)
        ntok = B * T
        predictor = {"ms": pred_ms, "tokens": ntok, "fp32_flops_per_token": fl_tok, "bytes_per_token": by_tok,
                     "achieved_tflops_fp32_equiv": ntok * fl_tok / (pred_ms * 1e-3) / 1e12,
                     "achieved_gbs": ntok * by_tok / (pred_ms * 1e-3) / 1e9,
                     "binds": "neither HBM nor FMA throughput: the BiLSTM's serial chain (T steps x n_lstm layers x ~1.7 us per "
                              "recurrent step: tcgen05 issue floor + gate math + DSMEM exchange of h_t across the 8-CTA cluster)",
                     "chain_us": T * nl * 1.7}
        # ---- roofline of the dominant kernel (tcgen05 GEMM family), timed in situ -------------
        path.set_option("profile", 1)
        for _ in range(2):
            dev_step()
        prof = path.profile_read()
        path.set_option("profile", 0)
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak, peak_src = (peaks["bf16_tflops"], "measured burst (MEASURED_PEAKS.json bf16_tflops)") \
            if "bf16_tflops" in peaks else (1590.0, "fallback (B200_PROFILING.md)")
        ms, flops, n = prof["gemm_tc"]
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "gemm_traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        # the denoiser's GEMM mix of one evaluation (x layers), M = 2*B*K rows: (name, N, K, epilogue, count)
        R, L, d, dff = 2 * B * cfg.n_style, cfg.n_layers, cfg.d_model, cfg.d_ff
        # epilogue 6 = the product form of the residual GEMMs (GEMM + gated residual + AdaLN in one kernel); only the
        # GEMM's 2MNK flops are credited to it
        mix = [("qkv", 3 * d, d, 2, L), ("attn_out+ln", d, d, 6, L), ("q_cross", d, d, 2, L), ("cross_out+ln", d, d, 6, L),
               ("ffn1_gelu", dff, d, 3, L), ("ffn2+ln", d, dff, 6, L)]
        per_shape, tot_flops, tot_us = {}, 0.0, 0.0
        for name, N, K, epi, cnt in mix:
            us = path.bench_gemm(R, N, K, epi, 50)
            fl = 2.0 * R * N * K
            per_shape[name] = {"M": R, "N": N, "K": K, "us": round(us, 2), "tflops": round(fl / us * 1e-6, 1)}
            tot_flops += fl * cnt
            tot_us += us * cnt
        ach = tot_flops / tot_us * 1e-6
        ach_situ = flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        tot_ms = sum(v[0] for v in prof.values())
        n_mix = sum(c for *_, c in mix)
        roofline = {"kernel": "gemm2_kernel / gemmln3_kernel (persistent tcgen05/TMEM/TMA bf16 GEMM family of the denoiser)", "bound": "tensor",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                    "peak_source": peak_src, "avg_launch_us": tot_us / n_mix, "flops_per_launch": tot_flops / n_mix,
                    "how": "per shape of one denoiser evaluation's GEMM mix (M = 2*B*K rows): 50 back-to-back launches of the "
                           "product kernel on the launching stream between two CUDA events (PDL-chained as in the "
                           "evaluation loop, L2-warm operands); achieved = sum(count * 2MNK) / sum(count * avg time)",
                    "per_shape": per_shape,
                    "in_situ_event_pairs": {"achieved": ach_situ, "frac": ach_situ / peak, "launches_per_step": n // 2,
                                            "avg_launch_us": ms / max(n, 1) * 1e3,
                                            "note": "eager profile mode, one CUDA-event pair per launch inside the real step: "
                                                    "includes event/launch gaps and loses PDL overlap (lower bound)"},
                    "share_of_profiled_step": ms / tot_ms if tot_ms > 0 else None,
                    "classes_ms_per_step": {k: v[0] / 2 for k, v in prof.items()}}
        if world == 1:
            r = run_oracle(B, 1, 1)
            cpu = {"value": r["utt_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                   "sample": f"one full cfg2 batch ({B} utterances) through the fp32 PyTorch oracle, {r['wall_s']:.1f} s wall"}
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
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "predictor": predictor,
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
    ap.add_argument("--ncu", action="store_true",
                    help="profiling run: cudaProfilerStart/Stop around the last timed step, skip the roofline/cpu legs "
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
