#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/configs.py config3 > gpurun_out/config3.jsonl 2>&1
timeout 600 python tools/configs.py config4 > gpurun_out/config4.jsonl 2>&1
timeout 600 python tools/configs.py fp32 > gpurun_out/config4_fp32.jsonl 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:actor_cost_tma -s 4 -c 1 -o gpurun_out/actor_tma_full \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_tma.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/config3.jsonl gpurun_out/config4.jsonl gpurun_out/config4_fp32.jsonl
