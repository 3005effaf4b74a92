set -u
mkdir -p gpurun_out
for mb in 1 5 6 8; do
  echo "== RCG_FITW_MINB=$mb"
  RCG_FITW_MINB=$mb RCG_PHASES=1 python tools/configs.py config3 --envs 131072 --t1 2.0 2>&1 | cut -c1-420
done
for mb in 1 6; do
  echo "== 1M RCG_FITW_MINB=$mb"
  RCG_FITW_MINB=$mb RCG_PHASES=1 python tools/configs.py config3 --envs 1048576 --t1 2.0 2>&1 | cut -c1-420
done
