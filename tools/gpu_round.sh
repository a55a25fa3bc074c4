#!/bin/bash
# One GPU-box session: tests, bench (both arms), ncu launch list, flop counters and a full capture of the top kernels.
set -x
O=gpurun_out/$1
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/pytest.log
python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2>> $O/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --roadmap-n 262144 > $O/bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_target.csv python tools/profile_target.py > $O/target.log 2>&1
ncu --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fma.sum,smsp__inst_executed.sum --clock-control none -k regex:validate_motions --csv --log-file $O/validate_flops.csv python tools/profile_target.py --no-roadmap > $O/target_flops.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:validate_motions -s 1 -c 2 -o $O/prof_validate python tools/profile_target.py --no-roadmap > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_cells -c 1 -o $O/prof_knn_cells python tools/profile_target.py > /dev/null 2>&1
ls -la $O
