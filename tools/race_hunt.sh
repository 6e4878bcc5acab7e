#!/bin/bash
# NOTE: since the end of round 2 the cluster tile scan is compiled only into the `cluster_scan` variant (python tools/build_variants.py
# cluster_scan; run with GSR_LIB=gaustudio_b200/variants/libgsr_cluster_scan.so): with the default library GSR_SCAN_CLUSTER is ignored.
for cl in 1 0; do
  GSR_SCAN_CLUSTER=$cl timeout -k 5 100 python tools/race_hunt.py 40000 40 2>&1 | grep -v Warning | tail -12
  GSR_SCAN_CLUSTER=$cl timeout -k 5 100 python tools/race_hunt.py 150000 30 2>&1 | grep -v Warning | tail -12
done
