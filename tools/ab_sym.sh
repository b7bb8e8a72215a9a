python -m pytest tests/test_gpu_float.py tests/test_gpu_baseline_shapes.py -x -q 2>&1 | tail -2
for lib in "" pqa2_b200/build/lib_sym0.so ""; do
  B200VMAF_LIB=$lib python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d.get('kernels',{})
        print('$lib', round(d['value'],1), {n: v['ms_per_launch'] for n,v in k.items() if v['ms_per_launch']>0.05})
"
done
