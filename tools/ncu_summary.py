"""Summarise an ncu --set full report (raw page csv) into the few numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'smsp__inst_executed_op_global_ld.sum', 'sm__sass_thread_inst_executed_op_fp32_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fp32_pred_on.sum']
want += sorted(h for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued'))
for w in want:
    if w in idx:
        print("%-72s %-8s %s" % (w, units[idx[w]], [r[idx[w]][:28] for r in rows[2:]]))
