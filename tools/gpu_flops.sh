#!/bin/bash
# ncu flop / traffic counters of the bench's all-valid launches (grid variant = launch #2, dense variant = launch #3) -> validate_flops.json
O=gpurun_out/$1
mkdir -p $O
ncu --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fma.sum,smsp__inst_executed.sum --clock-control none -k regex:validate_motions --csv --log-file $O/validate_flops.csv python tools/profile_target.py --no-roadmap > $O/target_flops.log 2>&1
tail -1 $O/target_flops.log > $O/facts.json
python tools/summarize_ncu.py flops $O/validate_flops.csv $O/facts.json > $O/validate_flops.json
cat $O/validate_flops.json
