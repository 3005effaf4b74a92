set -u
mkdir -p gpurun_out
for g in 0 -1 -2 -3 -4 -6 -8; do
  echo "== grid $g"; RCG_ACTOR_CTAS_PER_SM=$g python tools/exp_overlap.py --variants 1,2,2n,3,4 2>&1 | tail -n 6
done
