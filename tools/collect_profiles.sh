#!/bin/bash
# gpurun_out/<run> (tools/gpu_round3.sh) -> the tracked summaries profiles/<tag>_*. usage: tools/collect_profiles.sh <run> <tag>
set -e
cd "$(dirname "$0")/.."
S=gpurun_out/$1; T=profiles/$2
cp $S/bench_n1.json ${T}_bench_n1.json
cp $S/bench_reference.json ${T}_bench_reference_arm.json
cp $S/launches_bench.csv ${T}_launches_bench.csv
python tools/summarize_ncu.py launches $S/launches_bench.csv > ${T}_launch_shares.md
cp $S/validate_flops.csv ${T}_validate_flops.csv
cp $S/validate_flops.json profiles/validate_flops.json
cp $S/pytest.log ${T}_pytest_gpu.log
[ -f $S/prof_motions_allvalid.ncu-rep ] && python tools/ncu_key.py $S/prof_motions_allvalid.ncu-rep > ${T}_motions_allvalid_metrics.txt
[ -f $S/prof_rakes_roadmap.ncu-rep ] && python tools/ncu_key.py $S/prof_rakes_roadmap.ncu-rep > ${T}_rakes_roadmap_metrics.txt
[ -f $S/prof_rakes_roadmap.ncu-rep ] && python tools/ncu_regions.py $S/prof_rakes_roadmap.ncu-rep > ${T}_rakes_kernel_regions.txt
[ -f $S/prof_knn_point.ncu-rep ] && python tools/ncu_key.py $S/prof_knn_point.ncu-rep > ${T}_knn_point_metrics.txt
ls -la ${T}_*
