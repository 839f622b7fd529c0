#!/bin/bash
# one-screen summary of an .ncu-rep: tools/ncu_summary.sh rep.ncu-rep
ncu -i "$1" --page details 2>/dev/null | grep -E "Duration|Registers Per|Theoretical Occ|Achieved Occ|Executed Ipc Active|No Eligible|Eligible Warps|DRAM Throughput|L1/TEX Cache Thr|L2 Cache Thr|Compute \(SM\)|Issue Slots Busy|Shared Memory Config|Dynamic Shared"
ncu -i "$1" --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]; v=rows[2] if len(rows)>2 else rows[1]
want=['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__t_sectors_srcunit_tex_op_red.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__cycles_elapsed.avg','launch__grid_size']
for w in want:
    for i,c in enumerate(h):
        if c==w: print(w, v[i], rows[1][i])
st=[(float(v[i]),c) for i,c in enumerate(h) if c.startswith('smsp__average_warps_issue_stalled') and c.endswith('_per_issue_active.ratio') or c.startswith('smsp__average_warp_latency_issue_stalled')]
for x in sorted(st,reverse=True)[:8]: print(x)
"
