#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
B200MP_FIRST_LOOK=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/l_nofirst.csv python tools/profile_target.py --no-roadmap > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/l_first.csv python tools/profile_target.py --no-roadmap > /dev/null 2>&1
B200MP_MOTIONS_BUCKETED=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/l_plain.csv python tools/profile_target.py --no-roadmap > /dev/null 2>&1
for f in l_plain l_nofirst l_first; do echo == $f; python - <<PY
import csv
rows=[r for r in csv.DictReader([l for l in open("$O/$f.csv") if l.startswith('"')])]
for r in rows:
    print(r["ID"], r["Kernel Name"][:60].replace("b200mp::",""), r["Metric Value"], r.get("Metric Unit"))
PY
done
