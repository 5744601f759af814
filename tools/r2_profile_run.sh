#!/bin/bash
# Round-2 profiling pass on the GPU box (run under gpurun): the ncu launch list of a short bench.py run with DRAM traffic of
# every step kernel, and one full capture of the dominant kernel (a whole 131,072-world launch).
O=gpurun_out/r2prof; mkdir -p $O
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --t0-worlds 0 --no-invariance --ccd-steps 0"
# step-kernel launches before the timed steps: (settle 300 + warm-up 3) x 9 kernels (init_proxies does not match the filter)
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'(collide|plan|prepare|solve_velocity|solve_position|solve_position_joint|sync|broadphase|toi)_kernel_f32' \
    --launch-skip 2727 --launch-count 18 --csv --log-file $O/launches.csv $BENCH > $O/bench_under_ncu.json 2> $O/ncu.err
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:solve_velocity_kernel --launch-skip 303 --launch-count 1 \
    -o $O/solve_velocity $BENCH > /dev/null 2>> $O/ncu.err
ncu -i $O/solve_velocity.ncu-rep --page raw --csv > $O/solve_velocity_raw.csv 2>> $O/ncu.err
ncu -i $O/solve_velocity.ncu-rep --page source --csv --print-source sass > $O/solve_velocity_src.csv 2>> $O/ncu.err
python tools/ncu_summary.py $O/solve_velocity.ncu-rep > $O/solve_velocity_summary.txt 2>> $O/ncu.err
python tools/ncu_hot.py $O/solve_velocity_src.csv 30 > $O/solve_velocity_hot.txt 2>> $O/ncu.err
python tools/ncu_traffic.py $O/launches.csv pyramid 131072 > $O/traffic.txt 2>> $O/ncu.err
cp profiles/r2_traffic.json $O/
rm -f $O/solve_velocity.ncu-rep $O/solve_velocity_src.csv
ls -la $O; cat $O/solve_velocity_summary.txt; tail -5 $O/ncu.err
