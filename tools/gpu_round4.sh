#!/bin/bash
# Final 1-GPU session of the round: tools/gpu_round3.sh, then compute-sanitizer (memcheck, racecheck) over the paths this session changed
bash tools/gpu_round3.sh $1 > gpurun_out/$1.round3.log 2>&1
O=gpurun_out/$1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_validate_gpu.py tests/test_roadmap_gpu.py -m gpu -x -q -k "first_look_and_queue_takes and panda or matches_checker or c2_sample" > $O/memcheck.log 2>&1
echo "memcheck rc=$?" > $O/sanitizer_summary.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_roadmap_gpu.py -m gpu -x -q -k "matches_checker" > $O/racecheck.log 2>&1
echo "racecheck rc=$?" >> $O/sanitizer_summary.log
tail -3 $O/memcheck.log >> $O/sanitizer_summary.log; tail -3 $O/racecheck.log >> $O/sanitizer_summary.log
cat $O/pytest.log; cut -c1-400 $O/bench_n1.json; cat $O/sanitizer_summary.log
