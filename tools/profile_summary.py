"""profiles/ summaries from a gpurun round: python tools/profile_summary.py <tag>   (reads gpurun_out/*_<tag>*)"""
import csv, json, re, subprocess, sys, io, shutil

tag = sys.argv[1]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sectors.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']


def summarise(rep, dst):
    H, U, rows = raw(rep)
    idx = {w: H.index(w) for w in WANT if w in H}
    out = []
    for r in rows:
        e = {'kernel': re.sub(r'\(.*', '', r[H.index('Kernel Name')]).replace('void ', '').replace('stz::', '')}
        for w, i in idx.items():
            e[w] = r[i] + ' ' + U[i]
        out.append(e)
    json.dump(out, open(dst, 'w'), indent=1)
    return H, U, rows


H, U, rows = summarise(f"gpurun_out/prof_{tag}_layer.ncu-rep", "profiles/r01_ncu_full_layer_final.json")
num = lambda x: float(x.replace(',', ''))
mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
g = [r for r in rows if 'gemm2_kernel' in r[H.index('Kernel Name')] or 'gemmln3_kernel' in r[H.index('Kernel Name')]]
ir, iw = H.index('dram__bytes_read.sum'), H.index('dram__bytes_write.sum')
tr = [num(r[ir]) * mult[U[ir]] + num(r[iw]) * mult[U[iw]] for r in g]
json.dump({"kernel": "gemm2_kernel / gemmln3_kernel family (one denoiser layer of cfg2, cold-cache ncu --set full replay)", "launches": len(g),
           "dram_bytes_per_launch": sum(tr) / len(tr),
           "per_launch": [{"kernel": re.sub(r'\(.*', '', r[H.index('Kernel Name')]).replace('void ', ''), "dram_bytes": t} for r, t in zip(g, tr)],
           "source": "profiles/r01_ncu_full_layer_final.json (ncu --set full --clock-control none)"},
          open('profiles/gemm_traffic.json', 'w'), indent=1)
try:
    summarise(f"gpurun_out/prof_{tag}_lstm.ncu-rep", "profiles/r01_ncu_full_lstm_final.json")
except Exception as e:
    print("no lstm capture", e)
shutil.copy(f"gpurun_out/launches_{tag}.csv", "profiles/r01_launches_final.csv")
shutil.copy(f"gpurun_out/bench_{tag}.json", "profiles/r01_bench_final.json")
s = subprocess.run([sys.executable, "tools/agg_launches.py", "profiles/r01_launches_final.csv"], capture_output=True, text=True).stdout
open("profiles/r01_launches_final_summary.txt", "w").write(s)
print(s)
