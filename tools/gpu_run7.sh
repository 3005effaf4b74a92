#!/bin/bash
# 2-GPU validation of the sharded bench (NCCL gather/reduce at the end), both arms, launched like the driver does.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus2.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 300 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 100 --warmup 3 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "rc=$?" >> gpurun_out/bench_ref_n2.err
cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_ref_n2.json; tail -3 gpurun_out/bench_ref_n2.err
