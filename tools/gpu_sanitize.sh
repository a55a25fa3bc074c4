#!/bin/bash
# compute-sanitizer memcheck over a few small parity tests (grid + ordered variants, attachments, sessions, roadmap, kNN index)
O=gpurun_out/$1
mkdir -p $O
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_validate_gpu.py -m gpu -x -q -k "golden_validity and (panda_clutter or panda_cage or fetch_terrain) or golden_attachments and panda" > $O/memcheck_validate.log 2>&1
echo "validate rc=$?" > $O/summary.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_session_gpu.py -m gpu -x -q -k "restarts or snapshot" > $O/memcheck_session.log 2>&1
echo "session rc=$?" >> $O/summary.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_roadmap_gpu.py -m gpu -x -q -k "matches_checker or incremental" > $O/memcheck_roadmap.log 2>&1
echo "roadmap rc=$?" >> $O/summary.log
cat $O/summary.log; tail -4 $O/memcheck_validate.log; tail -4 $O/memcheck_session.log; tail -4 $O/memcheck_roadmap.log; grep -c "Invalid\|ERROR SUMMARY: [1-9]" $O/memcheck_*.log
