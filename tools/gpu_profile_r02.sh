#!/bin/bash
# ncu evidence of round 2 (run from the repo root under gpurun, one GPU): the launch list of the bench command, one
# `ncu --set full` capture of the dominant kernel (actor_cost_tma_kernel, as bench.py launches it: one of two environment
# blocks per launch) and one of the in-loop rk45_kernel<..., CTRL=1> at 65,536 lanes (bench.py --blocks 1), summarised to text.
set -u
mkdir -p gpurun_out
B="--steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-opt --no-extra --no-reference-python"
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py $B > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:actor_cost_tma_kernel -s 6 -c 1 -f -o /tmp/actor_tma \
    python bench.py $B > gpurun_out/ncu_actor.log 2>&1
python tools/ncu_summary.py /tmp/actor_tma.ncu-rep > gpurun_out/r02_actor_tma_ncu_full.txt 2>&1
python tools/ncu_hot.py /tmp/actor_tma.ncu-rep actor_cost_tma_kernel 1 1.0 > gpurun_out/r02_actor_tma_hot.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:rk45_kernel -s 8 -c 1 -f -o /tmp/rk45_adv \
    python bench.py $B --blocks 1 > gpurun_out/ncu_rk45.log 2>&1
python tools/ncu_summary.py /tmp/rk45_adv.ncu-rep > gpurun_out/r02_rk45_advance_inloop_ncu_full.txt 2>&1
python tools/ncu_hot.py /tmp/rk45_adv.ncu-rep rk45_kernel 1 1.0 > gpurun_out/r02_rk45_advance_inloop_hot.txt 2>&1
head -n 30 gpurun_out/r02_rk45_advance_inloop_ncu_full.txt
