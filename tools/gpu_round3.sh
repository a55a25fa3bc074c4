#!/bin/bash
# Full 1-GPU session: tests, ncu flop counters -> validate_flops.json, bench (both arms), ncu launch list, full captures of the top kernels.
O=gpurun_out/$1
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/pytest.log
bash tools/gpu_flops.sh $1 > /dev/null 2>&1
cp $O/validate_flops.json profiles/validate_flops.json
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2>> $O/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-configs --roadmap-n 262144 > $O/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:validate_motions -s 2 -c 1 -o $O/prof_motions_allvalid python tools/profile_target.py --no-roadmap > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:validate_rakes -s 8 -c 1 -o $O/prof_rakes_roadmap python tools/profile_target.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_point -c 1 -o $O/prof_knn_point python tools/profile_target.py > /dev/null 2>&1
cat $O/pytest.log; tail -c 1500 $O/bench_n1.json; echo; cat $O/bench_reference.json | cut -c1-300; cat $O/validate_flops.json | head -30; ls -la $O
