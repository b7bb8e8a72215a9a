#!/bin/bash
# Source-level ncu capture of one kernel: tools/prof_src.sh <workload> <kernel-regex> <count> <tag>
# Writes gpurun_out/src_<tag>.csv (per-line SASS/source counters) and gpurun_out/raw_<tag>.csv.
set -e
WL=$1; K=$2; N=$3; TAG=$4
ncu --set full --import-source on --clock-control none -k "regex:$K" -c "$N" -f -o gpurun_out/prof_$TAG \
    python bench.py --workload "$WL" --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_$TAG.log 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/src_$TAG.csv 2>/dev/null || \
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/src_$TAG.csv
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv
rm -f gpurun_out/prof_$TAG.ncu-rep
