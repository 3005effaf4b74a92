set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -n 40 > gpurun_out/t1.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
for g in 0 -1 -2 -4; do
  echo "== grid $g"; RCG_ACTOR_CTAS_PER_SM=$g python tools/exp_overlap.py --variants 1,2,2n,3,4 2>&1 | tail -n 6
done > gpurun_out/exp1.log 2>&1
echo "== default policy" >> gpurun_out/exp1.log; python tools/exp_overlap.py --variants 1,2,3,4 2>&1 | tail -n 5 >> gpurun_out/exp1.log
python bench.py --steps 200 --no-extra --no-opt --no-reference-python > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
python bench.py --steps 20 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
python bench.py --impl reference --steps 20 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -n 3 gpurun_out/t1.log; cat gpurun_out/smoke.log | tail -n 2; tail -c 600 gpurun_out/bench_a.err; tail -c 600 gpurun_out/bench_b.err
