#!/bin/bash
# End-of-round evidence run (1 x B200): ncu --set full of one steady-state step, the launch list of bench.py, then
# the bench lines (both arms, cfg5, training step).  Numbers printed under ncu are never quoted.
set -x
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_ -s 24 -c 12 -f -o gpurun_out/r2_prof_final2 python tools/profile_one.py cfg3 4 > gpurun_out/r2_prof_final2.log 2>&1
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final2_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --dropin 0 > gpurun_out/r2_final2_launches.log 2>&1
timeout -k 10 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_final2_bench_reference.json 2> gpurun_out/r2_final2_bench_reference.err
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final2_bench_new.json 2> gpurun_out/r2_final2_bench_new.err
timeout -k 10 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2_final2_bench_new_40.json 2> gpurun_out/r2_final2_bench_new_40.err
timeout -k 10 600 python bench.py --config cfg5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_final2_bench_cfg5.json 2> gpurun_out/r2_final2_bench_cfg5.err
timeout -k 10 600 python bench.py --train 1 --steps 20 --warmup 6 > gpurun_out/r2_final2_train_1gpu.json 2> gpurun_out/r2_final2_train_1gpu.err
tail -c 600 gpurun_out/r2_final2_bench_new.json

