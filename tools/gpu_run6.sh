#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/kernel_bench.py --what critic > gpurun_out/kb_critic.jsonl 2>&1
timeout 300 python tools/kernel_bench.py --what rk45 > gpurun_out/kb_rk45.jsonl 2>&1
timeout 900 python bench.py --steps 500 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
for c in 1 2 8; do timeout 300 python bench.py --steps 300 --e2e-chunks $c --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks',$c, d['e2e'])" >> gpurun_out/e2e_chunks.log 2>&1; done
tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/kb_critic.jsonl gpurun_out/kb_rk45.jsonl; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/e2e_chunks.log
