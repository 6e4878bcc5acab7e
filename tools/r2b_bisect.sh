#!/bin/bash
# NOTE: since the end of round 2 the cluster tile scan is compiled only into the `cluster_scan` variant (python tools/build_variants.py
# cluster_scan; run with GSR_LIB=gaustudio_b200/variants/libgsr_cluster_scan.so): with the default library GSR_SCAN_CLUSTER is ignored.
# which of the two late changes breaks test_fused_path_with_mostly_culled_ctas_matches_unfused?
mkdir -p gpurun_out
T=tests/test_gpu_api.py::test_fused_path_with_mostly_culled_ctas_matches_unfused
for spec in 1 0; do for cl in 1 0; do
  GSR_SPECULATE=$spec GSR_SCAN_CLUSTER=$cl timeout -k 10 120 python -m pytest $T -x -q 2>&1 | grep -E "AssertionError:|passed|failed" | tail -2 | sed "s/^/spec=$spec cluster=$cl: /"
done; done
GSR_SPECULATE=1 timeout -k 10 120 python -m pytest tests/test_gpu_api.py::test_exact_mode_speculation_is_invisible -x -q 2>&1 | tail -15
timeout -k 10 400 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest_all.txt 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2b_pytest_all.txt | tail -20
