set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_critic_refit.py tests/test_dropin_api.py -m gpu -q -s 2>&1 | grep -E "refit loop vs|passed|failed|FAILED|Error" | tail -n 20 > gpurun_out/t_refit.log
RCG_PHASES=1 python tools/configs.py config3 --envs 1048576 --t1 2.0 > gpurun_out/c3_1m_exact.jsonl 2>&1
RCG_PHASES=1 python tools/configs.py config3 --envs 131072 --t1 2.0 > gpurun_out/c3_131k_exact.jsonl 2>&1
python tools/kernel_bench.py --what opt > gpurun_out/kernel_bench_opt.jsonl 2>&1
python tools/kernel_bench.py --what critic > gpurun_out/kernel_bench_critic.jsonl 2>&1
cat gpurun_out/t_refit.log; cat gpurun_out/c3_1m_exact.jsonl gpurun_out/c3_131k_exact.jsonl | cut -c1-700; cut -c1-420 gpurun_out/kernel_bench_opt.jsonl
