#!/bin/bash
# usage: tools/ncu_rep_summary.sh <file.ncu-rep> "<title>"  — the recipe's metric grep (B200_PROFILING.md) as text
rep=$1; title=$2
echo "# ncu --set full --clock-control none --import-source on : $title"
echo "# source: $(basename $rep)"
ncu -i $rep --page raw --csv 2>/dev/null | python3 -c "
import sys, csv, re
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
want = re.compile(r'^(Kernel Name|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|dram__cycles_active\.avg\.pct|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|sm__pipe_tensor_cycles_active\.avg|sm__inst_executed_pipe_tensor_subpipe_hmma\.avg|sm__warps_active\.avg\.pct|launch__registers_per_thread$|launch__grid_size|launch__block_size|sm__throughput\.avg\.pct|lts__t_bytes\.sum$|smsp__inst_executed\.sum$|sm__cycles_elapsed\.max$|launch__shared_mem_per_block_dynamic|sm__inst_executed_pipe_xu\.avg|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$)')
for h, u, v in zip(hdr, units, vals):
    if want.search(h):
        print(f'{h:90s} {v} {u}')
"
