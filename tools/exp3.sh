set -u
mkdir -p gpurun_out
RCG_PHASES=1 python tools/configs.py config3 --envs 131072 --t1 2.0 > gpurun_out/c3_131k.jsonl 2>&1
RCG_PHASES=1 python tools/configs.py config3 --envs 1048576 --t1 2.0 > gpurun_out/c3_1m.jsonl 2>&1
python tools/configs.py config3 --envs 131072 --t1 2.0 >> gpurun_out/c3_131k.jsonl 2>&1
python tools/dump_critic_problems.py > gpurun_out/dump.log 2>&1
cat gpurun_out/c3_131k.jsonl gpurun_out/c3_1m.jsonl | cut -c1-900
