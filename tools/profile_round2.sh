#!/bin/bash
# Round-2 evidence run on the GPU box: the plain bench line first, then -- only after it exited -- the ncu launch list of
# the same command and one `--set full` capture of a whole frame group per workload.  Usage: tools/profile_round2.sh <tag>
TAG=${1:-v1}
O=gpurun_out
python bench.py --steps 20 --warmup 3 > $O/r02_bench_$TAG.json 2> $O/r02_bench_$TAG.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference_$TAG.json 2>> $O/r02_bench_$TAG.err
X="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --contexts 1"     # one context: the launches of a frame group stay in order
# numbers printed under ncu are never bench values
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02_launches_$TAG.csv \
    python bench.py $X > $O/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --launch-skip 1176 -c 26 -f -o $O/full_float_$TAG python bench.py $X > $O/ncu_full_float_$TAG.log 2>&1
ncu -i $O/full_float_$TAG.ncu-rep --page raw --csv > $O/raw_float_$TAG.csv
if [ -z "$FLOAT_ONLY" ]; then      # FLOAT_ONLY=1: only bv_float.cu changed since the last full run
ncu --set full --clock-control none --launch-skip 590 -c 14 -f -o $O/full_int_$TAG python bench.py --workload 1080p-int $X > $O/ncu_full_int_$TAG.log 2>&1
ncu -i $O/full_int_$TAG.ncu-rep --page raw --csv > $O/raw_int_$TAG.csv
ncu --set full --clock-control none --launch-skip 590 -c 14 -f -o $O/full_4k_$TAG python bench.py --workload 4k-int $X > $O/ncu_full_4k_$TAG.log 2>&1
ncu -i $O/full_4k_$TAG.ncu-rep --page raw --csv > $O/raw_4k_$TAG.csv
fi
rm -f $O/full_float_$TAG.ncu-rep $O/full_int_$TAG.ncu-rep $O/full_4k_$TAG.ncu-rep
ls -la $O | tail -12
# source-level capture of the two dominant kernels (dynamic SASS opcode histogram, per-line counters)
[ -z "$FLOAT_ONLY" ] && { tools/prof_src.sh 1080p-int "^vif_stat_kernel" 1 r02_int_vif0 || true; }
tools/prof_src.sh 1080p-float "^f_vif_stat_kernel" 1 r02_f_vif0 || true
[ -z "$FLOAT_ONLY" ] && { python tools/sass_hist.py $O/src_r02_int_vif0.csv 66355200 > $O/r02_sass_dyn_vif_stat_s0.txt 2>&1 || true; }
python tools/sass_hist.py $O/src_r02_f_vif0.csv 66355200 > $O/r02_sass_dyn_f_vif_stat_s0.txt 2>&1 || true
rm -f $O/src_r02_int_vif0.csv $O/src_r02_f_vif0.csv
ls -la $O | tail -8
