# Development aid: lanes per world of the TOI kernel (tools/build_variant.sh toiN -DB2B_TOI_LANES=N)
for v in toi4 toi16; do
  export B2B_LIB=$PWD/pypybox2d_b200/csrc/variants/libb2b_$v.so
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-invariance --e2e-steps 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['config']['continuous_on']['ms_per_step'], d['config']['from_t0']['ms_per_step'], d['config']['from_t0']['continuous_on']['ms_per_step'])"
done
