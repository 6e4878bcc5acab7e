#!/bin/bash
# compute-sanitizer passes over forward + backward (un-fused and fused) of a small scene (cfg1: 10k Gaussians, 256x256)
# driven straight through the binding (tools/sanitize_driver.py: no autograd or loss kernels in the process),
# once with the default build and once with the one-stage ring-stress build (tests/test_gpu_ring_stress.py).
# Usage (on a GPU box): bash tools/sanitize.sh  -> gpurun_out/r2_san_<tool>[_stress].txt
set -u
mkdir -p gpurun_out
for tool in memcheck synccheck initcheck racecheck; do
  for variant in "" "_stress"; do
    lib=""
    [ -n "$variant" ] && lib="$PWD/gaustudio_b200/libgsr_b200_stress.so"
    GSR_LIB="$lib" timeout -k 10 600 compute-sanitizer --tool $tool --print-limit 8 --error-exitcode 0 \
      python tools/sanitize_driver.py > gpurun_out/r2_san_${tool}${variant}.txt 2>&1
    echo "$tool$variant: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_san_${tool}${variant}.txt | tail -1)"
  done
done
