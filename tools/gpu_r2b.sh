#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
timeout 600 python -m pytest tests/test_knn_gpu.py tests/test_roadmap_gpu.py -m gpu -x -q 2>&1 | tail -8 > $O/pytest.log
timeout 900 python tools/knn_variants_probe.py '{}' '{"B200MP_KNN_STAT_PAIRS":1}' '{"B200MP_KNN_QPT":1}' '{"B200MP_KNN_QPT":1,"B200MP_KNN_STAT_PAIRS":1}' '{"B200MP_KNN_QPT":4}' '{"B200MP_KNN_QPT":4,"B200MP_KNN_STAT_PAIRS":1}' '{"B200MP_KNN_CELL_ROWS":32}' '{"B200MP_KNN_CELL_ROWS":128}' '{"B200MP_KNN_TAU0":0.2}' '{"B200MP_KNN_TAU0":0.5}' '{"B200MP_KNN_TAUG":1.5}' '{"B200MP_KNN_TAUG":3}' '{"B200MP_KNN_CHUNK":8}' '{"B200MP_KNN_CHUNK":128}' > $O/knn_variants.log 2>&1
cat $O/pytest.log $O/knn_variants.log
