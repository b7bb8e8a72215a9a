#!/bin/bash
# Round-end evidence run on the GPU box: tests, the three bench lines (+ reference arm), the ncu launch list and one
# `--set full` capture of a whole frame group per workload.  Usage: tools/profile_round.sh <tag>   (e.g. v4)
TAG=${1:-v4}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; tail -2 $O/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 3 > $O/bench_r01_1080p-float_$TAG.json 2> $O/bench_$TAG.err
python bench.py --workload 1080p-int --steps 20 --warmup 3 > $O/bench_r01_1080p-int_$TAG.json 2>> $O/bench_$TAG.err
python bench.py --workload 4k-int --steps 10 --warmup 3 > $O/bench_r01_4k-int_$TAG.json 2>> $O/bench_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_r01_reference_1080p-float_$TAG.json 2>> $O/bench_$TAG.err
# ncu passes: only after the plain runs above exited; numbers printed under ncu are never bench values
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_r01_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --launch-skip 500 -c 25 -f -o $O/full_float_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_full_float_$TAG.log 2>&1
ncu -i $O/full_float_$TAG.ncu-rep --page raw --csv > $O/raw_float_$TAG.csv
ncu --set full --clock-control none --launch-skip 280 -c 14 -f -o $O/full_int_$TAG \
    python bench.py --workload 1080p-int --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_full_int_$TAG.log 2>&1
ncu -i $O/full_int_$TAG.ncu-rep --page raw --csv > $O/raw_int_$TAG.csv
rm -f $O/full_float_$TAG.ncu-rep $O/full_int_$TAG.ncu-rep
ls -la $O | tail -15
