# Development aid: plan-kernel register caps (tools/build_variant.sh planNN -DB2B_PLAN_MIN_CTAS=NN) on the scenes that replan every step
mkdir -p gpurun_out/r3p
for v in base; do
  if [ $v = base ]; then unset B2B_LIB; else export B2B_LIB=$PWD/pypybox2d_b200/csrc/variants/libb2b_$v.so; fi
  for s in vertical_stack deep_stacks mixed; do
    w=""; [ $s = mixed ] && w="--worlds 32768"
    python bench.py --scene $s $w --steps 10 --warmup 3 --no-cpu-baseline --no-invariance --ccd-steps 0 --e2e-steps 1 > gpurun_out/r3p/${v}_$s.json 2> gpurun_out/r3p/${v}_$s.err
    python -c "
import json
d=json.loads(open('gpurun_out/r3p/${v}_$s.json').read().strip().splitlines()[-1])
print('$v $s', round(d['ms_per_step'],2), {k: round(x,2) for k,x in d['roofline']['kernel_ms'].items()})
"
  done
done
