set -u
mkdir -p gpurun_out
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N tools/kernel_bench.py --what sweep_multi > gpurun_out/sweep_multi_n$N.jsonl 2> gpurun_out/sweep_multi_n$N.err
done
wc -l gpurun_out/sweep_multi_n*.jsonl; tail -n 2 gpurun_out/sweep_multi_n8.jsonl | cut -c1-400
