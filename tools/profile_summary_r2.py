"""profiles/ summaries of a round-2 ncu pass: python tools/profile_summary_r2.py <tag>
reads gpurun_out/launches_<tag>.csv, prof_<tag>_layer.ncu-rep, prof_<tag>_pred.ncu-rep (tools/gpu_profile_r2.sh) and writes
profiles/r02_launches.csv, r02_launches_summary.txt, r02_ncu_full_layer.json, r02_ncu_predictor.json and
r02_roofline_evidence.json (what bench.py attaches to its roofline / predictor objects)."""
import collections, csv, io, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tag = sys.argv[1]
B, K, d, dff, L, E = 64, 50, 512, 2048, 8, 4          # cfg2
R = 2 * B * K


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sectors.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
MULT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}
num = lambda x: float(x.replace(',', ''))
kname = lambda s: re.sub(r'\(.*', '', s).replace('void ', '').replace('stz::', '')


def summarise(rep, dst):
    H, U, rows = raw(rep)
    idx = {w: H.index(w) for w in WANT if w in H}
    out = []
    for r in rows:
        e = {'kernel': kname(r[H.index('Kernel Name')]), 'grid': r[H.index('Grid Size')] if 'Grid Size' in H else None}
        for w, i in idx.items():
            e[w] = r[i] + ' ' + U[i]
        out.append(e)
    json.dump(out, open(dst, 'w'), indent=1)
    return H, U, rows


def val(H, U, r, name):
    i = H.index(name)
    return num(r[i]) * MULT.get(U[i], 1.0)


# ---- launch list ------------------------------------------------------------------------------------------------------
src = f"gpurun_out/launches_{tag}.csv"
shutil.copy(src, "profiles/r02_launches.csv")
rows = list(csv.reader(open(src)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
Hh, data = rows[hdr], rows[hdr + 1:]
ki, vi, ui = Hh.index('Kernel Name'), Hh.index('Metric Value'), Hh.index('Metric Unit')
agg, tot = collections.OrderedDict(), 0.0
for r in data:
    if len(r) <= vi:
        continue
    v = num(r[vi]) * MULT.get(r[ui], 1.0)
    a = agg.setdefault(kname(r[ki]), [0, 0.0]); a[0] += 1; a[1] += v; tot += v
lines = [f"{t:10.1f} us {n:5d} launches {t / n:8.2f} us/launch {100 * t / tot:5.1f}%  {k[:100]}" for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])]
lines.append(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches (ncu per-launch times are cold-cache and serialised: compare shares)")
open("profiles/r02_launches_summary.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
# the denoiser's layer GEMMs in that list: every gemmln3 launch + the gemm2 launches with the layer shapes' grids
layer_gemm_flops = E * L * 2.0 * R * d * (3 * d + d + d + d + dff + dff)
gemm_us = sum(t for k, (n, t) in agg.items() if k.startswith('gemmln3_kernel<0>') or k.startswith('gemm2_kernel<256, 3') or
              k.startswith('gemm2_kernel<192, 2') or k.startswith('gemm2_kernel<256, 2'))
all_gemm_us = sum(t for k, (n, t) in agg.items() if k.startswith('gemm'))

# ---- one denoiser layer, --set full -------------------------------------------------------------------------------------
H, U, rows = summarise(f"gpurun_out/prof_{tag}_layer.ncu-rep", "profiles/r02_ncu_full_layer.json")
g = [r for r in rows if 'gemm' in r[H.index('Kernel Name')]]
tw = sum(val(H, U, r, 'gpu__time_duration.sum') * num(r[H.index('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed')]) for r in g) / \
    sum(val(H, U, r, 'gpu__time_duration.sum') for r in g)
dram = [val(H, U, r, 'dram__bytes_read.sum') + val(H, U, r, 'dram__bytes_write.sum') for r in g]
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
ev = {"source": "tools/gpu_profile_r2.sh + tools/profile_summary_r2.py (ncu 2025.x, --clock-control none; per-launch times are cold-cache, serialised)",
      "launch_list": {"file": "profiles/r02_launches_summary.txt", "total_us": tot, "gemm_family_us": all_gemm_us,
                      "gemm_family_share": all_gemm_us / tot, "layer_gemm_us": gemm_us, "layer_gemm_tflop": layer_gemm_flops / 1e12,
                      "layer_gemm_tflops": layer_gemm_flops / gemm_us * 1e-6,
                      "layer_gemm_frac_of_burst": layer_gemm_flops / gemm_us * 1e-6 / peaks["bf16_tflops"]},
      "tensor_pipe_pct": {"time_weighted_over_layer_gemms": tw, "metric": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                          "file": "profiles/r02_ncu_full_layer.json", "launches": len(g)},
      "gemm_dram_bytes_per_launch": sum(dram) / len(dram),
      "gemm_dram_bytes": [{"kernel": kname(r[H.index('Kernel Name')]), "dram_bytes": t} for r, t in zip(g, dram)]}
json.dump(ev, open("profiles/r02_roofline_evidence.json", "w"), indent=1)
print(json.dumps({k: ev[k] for k in ("launch_list", "tensor_pipe_pct", "gemm_dram_bytes_per_launch")}, indent=1))

# ---- the predictor's and the elementwise kernels, --set full --------------------------------------------------------------
H, U, rows = summarise(f"gpurun_out/prof_{tag}_pred.ncu-rep", "profiles/r02_ncu_predictor_raw.json")
per = collections.OrderedDict()
for r in rows:
    k = kname(r[H.index('Kernel Name')])
    e = per.setdefault(k, {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
    e["launches"] += 1
    e["us"] += val(H, U, r, 'gpu__time_duration.sum')
    e["dram_bytes"] += val(H, U, r, 'dram__bytes_read.sum') + val(H, U, r, 'dram__bytes_write.sum')
for k, e in per.items():
    e["dram_gbs"] = e["dram_bytes"] / (e["us"] * 1e-6) / 1e9 if e["us"] else None
    e["dram_bytes_per_launch"] = e["dram_bytes"] / e["launches"]
    e["us_per_launch"] = e["us"] / e["launches"]
json.dump({"source": "ncu --set full, one cfg2 step (cold-cache replay: every first touch comes from HBM)", "hbm_peak_gbs": peaks["hbm_gbs"],
           "kernels": per}, open("profiles/r02_ncu_predictor.json", "w"), indent=1)
print(json.dumps(per, indent=1))
