"""Prints a compact set of ncu raw metrics + stall-reason breakdown from `ncu --page raw --csv` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
H, U = rows[0], rows[1]
for V in rows[2:]:
    D = {h: (U[i], V[i]) for i, h in enumerate(H)}
    print("==", D.get("Kernel Name", ("", ""))[1][:60])
    keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'launch__grid_size', 'launch__registers_per_thread',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
            'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed' if False else 'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
            'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
            'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
            'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
    for k in keys:
        if k in D:
            print(f"  {k:95s} {D[k][1]:>16s} {D[k][0]}")
    st = [(h, float(D[h][1])) for h in D if h.startswith('smsp__pcsamp_warps_issue_stalled') and D[h][1] not in ('', 'n/a')]
    tot = sum(v for _, v in st) or 1
    print("  -- stall reasons (pc sampling)")
    for h, v in sorted(st, key=lambda x: -x[1])[:10]:
        print(f"     {h[33:]:40s} {v:12.0f} {v / tot:6.3f}")
