set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -n 8 > gpurun_out/t_all.log
python tools/configs.py fp32 > gpurun_out/config4_fp32_report.jsonl 2> gpurun_out/config4_fp32.err
python tools/configs.py config4 > gpurun_out/config4.jsonl 2>&1
python tools/kernel_bench.py --what sweep > gpurun_out/sweep_config5.jsonl 2>&1
python tools/kernel_bench.py --what sweep_multi > gpurun_out/sweep_multi_n1.jsonl 2>&1
tail -n 4 gpurun_out/t_all.log; tail -c 1500 gpurun_out/config4_fp32_report.jsonl; tail -c 600 gpurun_out/config4_fp32.err; tail -n 2 gpurun_out/config4.jsonl | cut -c1-600; wc -l gpurun_out/sweep_config5.jsonl gpurun_out/sweep_multi_n1.jsonl
