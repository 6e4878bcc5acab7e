#!/bin/bash
# memcheck + synccheck of the default build on a dense scene (150k Gaussians on 320x320: ~1-3k instances per tile, so
# the 128-thread and the 512-thread sort tiers and multi-batch compositing lists all run) -- the quick re-check after a
# kernel change; tools/sanitize.sh is the full four-tool pass.
set -u
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  SAN_SCENE="cfg2,150000,320,320" timeout -k 10 900 compute-sanitizer --tool $tool --print-limit 8 --error-exitcode 0 \
    python tools/sanitize_driver.py > gpurun_out/r2_sanq_${tool}.txt 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY' gpurun_out/r2_sanq_${tool}.txt | tail -1) $(grep '^done' gpurun_out/r2_sanq_${tool}.txt)"
done
