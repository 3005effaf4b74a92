#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/kernel_bench.py --what headline > gpurun_out/kb_headline.jsonl 2>&1
RCG_ACTOR_NO_TMA=1 timeout 300 python tools/kernel_bench.py --what headline > gpurun_out/kb_headline_notma.jsonl 2>&1
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/kb_headline.jsonl; echo; cat gpurun_out/kb_headline_notma.jsonl; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:actor_cost_tma -s 4 -c 1 -o gpurun_out/actor_tma_full \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_tma.log 2>&1
