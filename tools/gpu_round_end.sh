#!/bin/bash
# Reproduces the measurements under profiles/ on a B200 box (run from the repo root, e.g.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round_end.sh'):
# parity tests, the kernel micro-benchmarks, configs 2/3/4, both bench arms, the ncu launch list of the bench command and
# one `ncu --set full` capture of the dominant kernel, summarised to text (the .ncu-rep files are too large to keep).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3 | tee gpurun_out/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log
for what in headline rk45 critic opt sweep; do
  python tools/kernel_bench.py --what $what > gpurun_out/kernel_bench_$what.jsonl 2>&1
done
python tools/configs.py config2 --t1 3.0 > gpurun_out/config2_actor_standins.jsonl 2>&1
python tools/configs.py config3 > gpurun_out/config3.jsonl 2>&1
python tools/configs.py config4 > gpurun_out/config4.jsonl 2>&1
python tools/configs.py fp32 > gpurun_out/config4_fp32_report.jsonl 2>&1
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-opt > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:actor_cost_tma_kernel -s 5 -c 1 -f -o /tmp/actor_tma \
    python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-opt > gpurun_out/ncu_actor.log 2>&1
python tools/ncu_summary.py /tmp/actor_tma.ncu-rep > gpurun_out/actor_tma_ncu_full.txt 2>&1
python tools/ncu_hot.py /tmp/actor_tma.ncu-rep actor_cost_tma_kernel 1 1.0 > gpurun_out/actor_tma_hot.txt 2>&1
cat gpurun_out/bench.json
