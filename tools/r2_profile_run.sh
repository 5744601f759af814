#!/bin/bash
# Round-2 profiling pass on the GPU box (run under gpurun): phase cycles of the non-settled scenes, the ncu launch list of a
# short bench.py run, DRAM traffic of every step kernel, and one full capture of the dominant kernel.
O=gpurun_out/r2e; mkdir -p $O
for sc in "deep_stacks 5 32768" "vertical_stack 2 65536" "mixed 4 16384"; do set -- $sc; timeout 300 python tools/gpu_scene_profile.py $1 $2 $3 > $O/phase_$1.txt 2>&1; done
export B2B_STEP_STREAMS=1
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --t0-worlds 0 --no-invariance"
# b2b launches before the timed steps: init_proxies 1 + settle 300 x 8 + warm-up 3 x 8
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:b2b:: \
    --launch-skip 2425 --launch-count 48 --csv --log-file $O/launches.csv $BENCH > $O/bench_under_ncu.json 2> $O/ncu.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_velocity_kernel --launch-skip 303 --launch-count 1 \
    -o $O/solve_velocity $BENCH > /dev/null 2>> $O/ncu.err
ncu -i $O/solve_velocity.ncu-rep --page raw --csv > $O/solve_velocity_raw.csv 2>> $O/ncu.err
ls -la $O
