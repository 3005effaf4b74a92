#!/bin/bash
# compute-sanitizer over the kernels that are new or changed in round 2 (run under gpurun, one GPU).
set -u
mkdir -p gpurun_out
out=gpurun_out/r02_sanitizer.txt
echo "compute-sanitizer (B200, round-2 code)" > $out
run() {  # tool, description, pytest args...
  tool=$1; shift; desc=$1; shift
  echo "--tool $tool  $desc:" >> $out
  timeout 900 compute-sanitizer --tool $tool python -m pytest "$@" -m gpu -q -x 2>&1 | grep -E "passed|failed|ERROR SUMMARY|error" | tail -n 4 >> $out
  echo "rc=$?" >> $out
}
run memcheck "disturbance lanes (rk45 DIST instantiation, rhs_disturbed, normal stream, engine)" tests/test_disturbance.py
run memcheck "critic fit with the exact line search (warp-per-environment kernel), recorded-weights loops, fp32 critic cost" tests/test_gpu_critic_refit.py tests/test_gpu_parity.py -k "refit or recorded or critic_fit or critic_cost_f32"
run memcheck "pipelined / host-staged loops, TMA kernel with the non-persistent grid, tensor-map cache" tests/test_gpu_parity.py -k "pipelined or staged or tma_kernel or last_actor_kernel or empty_batch"
run racecheck "TMA-staged actor kernel (shared-memory ring) under the pipelined loop" tests/test_gpu_parity.py -k "pipelined or tma_kernel"
cat $out
