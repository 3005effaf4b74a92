#!/bin/bash
# bench.py at N = 2, 4, 8 with the driver's arguments (run under `gpurun --gpus 8`); N = 1 comes from a 1-GPU call.
set -u
mkdir -p gpurun_out
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --steps 20 --warmup 5 --no-extra --no-opt > gpurun_out/scale_n8_b.json 2> gpurun_out/scale_n8_b.err
python - <<'PY'
import json
for f in ("scale_n2","scale_n4","scale_n8","scale_n8_b"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        x=d.get("extra") or {}
        print(f, "value %.4g ms %.4f e2e %.4g (%.4f)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]),
              ("c3 %.0f ms envsteps %.3g" % (x["config3_strong"]["ms_total"], x["env_steps"]["nominal"]["env_steps_per_s"])) if x else "")
    except Exception as e:
        print(f, "ERR", e)
PY
