#!/bin/bash
# A/B of CTA sizes of the memory-/latency-bound kernels (tools/build_variants.py builds the variant libraries;
# GSR_LIB selects one), plus the compositing experiments: GSR_RENDER_PAD caps the compositing CTAs per SM with a
# shared-memory pad, bwd48 caps the compositing backward at 48 registers.
# bench.py default mode (3 streams, one CUDA graph per view), alternating runs.
V=gaustudio_b200/variants
run() {  # name, lib ('' = default), pad
  GSR_LIB=${2:+$PWD/$V/libgsr_$2.so} GSR_RENDER_PAD=$3 timeout -k 10 300 python bench.py $BENCH_ARGS --steps ${STEPS:-60} --warmup 5 --no-cpu-baseline --dropin 0 > gpurun_out/r2_blk_$1.json 2>gpurun_out/r2_blk_$1.err
}
for i in 1 2; do
  run base_$i "" 0
  for v in "$@"; do run ${v}_$i $v 0; done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_blk_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = d["stages"]["_kernels_ms"]
        print(f.split("r2_blk_")[1][:-5].ljust(18), round(d["value"], 1), round(d["e2e"]["value"], 1),
              {n: k[n] for n in ("preprocess_fwd", "scatter", "tile_sort", "render_fwd", "render_bwd", "preprocess_bwd")})
    except Exception as e:
        print(f, "ERR", e)
PY
