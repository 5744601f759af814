"""Print the headline metrics and stall breakdown of an .ncu-rep (reads `ncu --page raw --csv`)."""
import csv, subprocess, sys
def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    return {k: (v[i], u[i]) for i, k in enumerate(h)}
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
if __name__ == "__main__":
    m = load(sys.argv[1])
    for k in WANT:
        if k in m: print("%-70s %s %s" % (k, m[k][0], m[k][1]))
    st = [(float(v[0].replace(',', '')), k) for k, v in m.items()
          if k.startswith('smsp__average_warp') and 'issue_stalled' in k and k.endswith('.ratio') and 'not_issued' not in k]
    for a, b in sorted(st)[::-1][:10]:
        print("%8.3f %s" % (a, b.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
