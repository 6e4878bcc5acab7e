#!/bin/bash
# A/B of the launch priorities (GSR_PRIORITY=1: memory-/latency-bound kernels at the device's highest priority, compositing
# kernels at the default one; 0: everything at the default), bench.py default mode, alternating runs.
for i in 1 2; do
  for p in 1 0; do
    GSR_PRIORITY=$p timeout -k 10 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --dropin 0 > gpurun_out/r2_prio${p}_$i.json 2>/dev/null
  done
done
for p in 1 0; do
  GSR_PRIORITY=$p timeout -k 10 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --dropin 0 --streams 4 > gpurun_out/r2_prio${p}_s4.json 2>/dev/null
  GSR_PRIORITY=$p GSR_TILE_ORDER=1 timeout -k 10 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --dropin 0 > gpurun_out/r2_prio${p}_lpt.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_prio*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1))
    except Exception as e:
        print(f, "ERR", e)
PY
