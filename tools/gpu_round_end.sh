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
python tools/configs.py config3 --t1 2.0 > gpurun_out/config3.jsonl 2>&1          # SURVEY 8d episode lengths: t1 = 2 / t1 = 100
python tools/configs.py config4 > gpurun_out/config4.jsonl 2>&1
python tools/exp_overlap.py --variants 1,2,3,4 > gpurun_out/overlap_blocks.jsonl 2>&1   # pipelined loop: blocks x stagger
python tools/configs.py fp32 > gpurun_out/config4_fp32_report.jsonl 2>&1
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
bash tools/gpu_profile_r02.sh            # ncu: launch list, actor TMA kernel, in-loop rk45_kernel<CTRL=1> (round 2)
cat gpurun_out/bench.json
