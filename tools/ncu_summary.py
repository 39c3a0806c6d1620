"""Print key metrics of every kernel in an .ncu-rep (raw page)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U = rows[0], rows[1]
want = ['Kernel Name', 'launch__grid_size', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum ', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma', 'smsp__issue_active.avg.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ', 'sm__pipe_fma_cycles_active', 'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum ', 'lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_ld.sum']
for r in rows[2:]:
    print('-----')
    for w in want:
        for i, h in enumerate(H):
            if h == w.strip() or (not w.endswith(' ') and h.startswith(w)):
                print(f"  {h} [{U[i]}] = {r[i][:110]}")
