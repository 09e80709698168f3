#!/bin/bash
# ncu counters of the sweep kernel with plain (0) and compact (1) nodes -> gpurun_out/cmp_{0,1}.csv (profiles/r2_compact_nodes_ncu.txt)
for c in 0 1; do
YART_TUNE_COMPACT=$c ncu --clock-control none -k regex:k_traverse --launch-skip 0 --launch-count 6 --csv --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,lts__t_sectors_op_read.sum,l1tex__t_sector_hit_rate.pct --log-file gpurun_out/cmp_$c.csv python tools/sweep.py --sets uniform --reps 1 > gpurun_out/cmp_$c.log 2>&1
done
