#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin_api.py -m gpu -q -x > gpurun_out/pytest_dropin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_dropin.log
timeout 300 python tools/kernel_bench.py --what headline > gpurun_out/kb_headline.jsonl 2>&1
timeout 300 python tools/kernel_bench.py --what rk45 > gpurun_out/kb_rk45.jsonl 2>&1
tail -40 gpurun_out/pytest_dropin.log; cat gpurun_out/kb_headline.jsonl gpurun_out/kb_rk45.jsonl
