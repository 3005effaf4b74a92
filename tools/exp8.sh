set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_critic_refit.py -m gpu -q 2>&1 | tail -n 3
# in-loop critic fit of config 3 (131,072 environments, an interval late in the episode): warp-per-environment phase
ncu --set full --clock-control none --import-source on -k regex:critic_fit3_warp_kernel -s 120 -c 1 -f -o /tmp/fitw \
    python tools/configs.py config3 --envs 131072 --t1 1.6 > gpurun_out/ncu_fitw.log 2>&1
python tools/ncu_summary.py /tmp/fitw.ncu-rep > gpurun_out/r02_critic_fit3_warp_exact_ncu_full.txt 2>&1
python tools/ncu_hot.py /tmp/fitw.ncu-rep critic_fit3_warp_kernel 1 1.5 > gpurun_out/r02_critic_fit3_warp_exact_hot.txt 2>&1
ncu --set full --clock-control none -k regex:critic_fit3_kernel -s 120 -c 1 -f -o /tmp/fit1 \
    python tools/configs.py config3 --envs 131072 --t1 1.6 > gpurun_out/ncu_fit1.log 2>&1
python tools/ncu_summary.py /tmp/fit1.ncu-rep > gpurun_out/r02_critic_fit3_phase1_ncu_full.txt 2>&1
grep -E "time_duration|inst_executed.sum|issue_active|pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active|warps_active|thread_inst_executed_per" gpurun_out/r02_critic_fit3_warp_exact_ncu_full.txt gpurun_out/r02_critic_fit3_phase1_ncu_full.txt
head -n 4 gpurun_out/r02_critic_fit3_warp_exact_hot.txt | cut -c1-700
