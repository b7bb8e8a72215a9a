#!/bin/bash
# A/B of library builds on one box: tools/ab_libs.sh <workload> <lib|""> ...   ("" = the in-tree library)
WL=$1; shift
for lib in "$@"; do
  B200VMAF_LIB=$lib python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d.get('kernels',{})
        print('[$lib]', round(d['value'],1), {n: round(v['ms_per_launch'],4) for n,v in k.items() if 'adm' in n})
"
done
