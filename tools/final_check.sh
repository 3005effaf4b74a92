set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -n 4 > gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -n 3 gpurun_out/t_all.log; tail -n 1 gpurun_out/smoke.log; tail -c 300 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_final.json")); r=json.load(open("gpurun_out/bench_ref_final.json"))
print("value %.4g ms %.4f e2e %.4g kernel %s frac %.3f frac_step %.3f same_config %s ratio e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["frac_step"], d["config"]==r["config"], d["e2e"]["value"]/r["value"]))
PY
