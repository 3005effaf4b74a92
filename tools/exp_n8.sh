set -u
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 200 --warmup 5 --no-extra --no-opt > gpurun_out/bench_n${N}_k200.json 2> gpurun_out/bench_n${N}_k200.err
tail -c 400 gpurun_out/bench_n$N.err; head -c 600 gpurun_out/bench_n$N.json
