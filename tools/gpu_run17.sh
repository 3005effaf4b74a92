#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/e2e_graph.log
for c in 2 4 8 16; do for g in "" "--no-graph"; do timeout 300 python bench.py --steps 300 --e2e-chunks $c $g --no-cpu-baseline 2>gpurun_out/e2e_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks',$c,'$g', d['e2e']['ms_per_step'], d['e2e']['value'])" >> gpurun_out/e2e_graph.log 2>&1; done; done
cat gpurun_out/e2e_graph.log; tail -5 gpurun_out/e2e_err.log
