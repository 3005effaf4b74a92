#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/kernel_bench.py --what critic > gpurun_out/kb_critic.jsonl 2>&1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 200 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk45_kernel -s 10 -c 2 -o gpurun_out/rk45_1m_full \
    python tools/kernel_bench.py --what rk45ni > gpurun_out/ncu_rk45_1m.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/kb_critic.jsonl; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_ref.json
