mkdir -p gpurun_out/r3v
for v in base csort; do
  if [ $v = base ]; then unset B2B_LIB; else export B2B_LIB=$PWD/pypybox2d_b200/csrc/variants/libb2b_$v.so; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-invariance --t0-worlds 0 --ccd-steps 0 > gpurun_out/r3v/$v.json 2> gpurun_out/r3v/$v.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r3v/$v.json').read().strip().splitlines()[-1])
print('$v', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['roofline']['kernel_ms'].items()})
"
done
