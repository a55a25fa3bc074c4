#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
timeout 900 python tools/knn_variants_probe.py '{"B200MP_KNN_QPT":1,"B200MP_KNN_CHUNK":8}' '{"B200MP_KNN_QPT":1,"B200MP_KNN_CHUNK":16}' '{"B200MP_KNN_QPT":1,"B200MP_KNN_CHUNK":4}' '{"B200MP_KNN_QPT":1,"B200MP_KNN_CHUNK":8,"B200MP_KNN_TAU0":0.4,"B200MP_KNN_TAUG":100}' '{"B200MP_KNN_QPT":1,"B200MP_KNN_CHUNK":8,"B200MP_KNN_TAU0":0.3,"B200MP_KNN_TAUG":100,"B200MP_KNN_STAT_PAIRS":1}' '{"B200MP_KNN_QPT":1,"B200MP_KNN_CHUNK":8,"B200MP_KNN_CELL_ROWS":32}' '{"B200MP_KNN_QPT":1,"B200MP_KNN_CHUNK":8,"B200MP_KNN_CELL_ROWS":32,"B200MP_KNN_STAT_PAIRS":1}' > $O/knn_variants.log 2>&1
export B200MP_KNN_QPT=1 B200MP_KNN_CHUNK=8
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/knn_launches.csv python tools/knn_ncu_target.py > $O/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_point -s 1 -c 1 -o $O/prof_knn_point python tools/knn_ncu_target.py > $O/ncu2.log 2>&1
cat $O/knn_variants.log; ls -la $O
