#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus8.txt
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 300 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?" >> gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --impl reference --gpus $N --steps 60 --warmup 3 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "rc=$?" >> gpurun_out/bench_ref_n$N.err
cat gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_ref_n$N.json; tail -2 gpurun_out/bench_ref_n$N.err
