#!/bin/bash
O=gpurun_out; TAG=v9
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_r01_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --launch-skip 500 -c 25 -f -o $O/full_float_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_full_float_$TAG.log 2>&1
ncu -i $O/full_float_$TAG.ncu-rep --page raw --csv > $O/raw_float_$TAG.csv
rm -f $O/full_float_$TAG.ncu-rep
cp $O/raw_int_v7.csv $O/raw_int_$TAG.csv 2>/dev/null
ls -la $O | grep v9
