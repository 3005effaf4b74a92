#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/kernel_bench.py --what headline > gpurun_out/kb_headline.jsonl 2>&1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -12 gpurun_out/pytest_gpu.log; cat gpurun_out/kb_headline.jsonl; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
