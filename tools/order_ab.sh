#!/bin/bash
# A/B of the CTA -> tile order of the one-CTA-per-tile kernels: GSR_TILE_ORDER=1 (longest first) vs 0 (raster), bench.py
# default mode (3 streams + CUDA graphs) and 4 streams, alternating runs.  -> gpurun_out/r2_order*.json
for i in 1 2; do
  for o in ${ORDERS:-1 0}; do
    GSR_TILE_ORDER=$o timeout -k 10 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --dropin 0 > gpurun_out/r2_order${o}_$i.json 2>/dev/null
  done
done
for o in ${ORDERS:-1 0}; do
  GSR_TILE_ORDER=$o timeout -k 10 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --dropin 0 --streams 4 > gpurun_out/r2_order${o}_s4.json 2>/dev/null
  GSR_TILE_ORDER=$o timeout -k 10 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --dropin 1 --streams 1 --graph 0 > gpurun_out/r2_order${o}_s1.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_order*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = d["stages"]["_kernels_ms"]
        print(f, round(d["value"], 1), round(d["e2e"]["value"], 1), k["render_fwd"], k["render_bwd"], k["tile_sort"], (d.get("dropin") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
