set -u
mkdir -p gpurun_out
python -m pytest tests/test_disturbance.py -m gpu -x -q 2>&1 | tail -n 30 > gpurun_out/t_dist.log
python -m pytest tests -m gpu -q 2>&1 | tail -n 12 > gpurun_out/t_all.log
tail -n 12 gpurun_out/t_dist.log; tail -n 6 gpurun_out/t_all.log
