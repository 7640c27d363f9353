"""Summarise an .ncu-rep (every profiled launch): the metrics the DESIGN/roofline discussion uses, pipe utilisation, stall reasons.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x.txt"""
import csv, sys, subprocess
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg', 'lts__t_requests_srcunit_tex_op_red.sum', 'lts__t_requests_srcunit_tex.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum', 'smsp__inst_executed_op_shared_atom.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_op_global_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum']
for vals in rows[2:]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print("==", d.get('Kernel Name', ('?',))[0])
    for h in want:
        if h in d:
            print(f"{h:80s} {d[h][0]:>18s} {d[h][1]}")
    for h, (v, u) in d.items():
        if ('sm__inst_executed_pipe' in h or 'sm__pipe' in h) and h.endswith('avg.pct_of_peak_sustained_active'):
            try:
                if float(v) > 5:
                    print(f"{h:95s} {float(v):.1f}")
            except ValueError:
                pass
    for h, (v, u) in d.items():
        if 'smsp__average_warp' in h and 'issue_stalled' in h and 'ratio' in h:
            try:
                if float(v) > 0.25:
                    print(f"{h:95s} {float(v):.2f}")
            except ValueError:
                pass
    print()
