#!/bin/bash
# NOTE: since the end of round 2 the cluster tile scan is compiled only into the `cluster_scan` variant (python tools/build_variants.py
# cluster_scan; run with GSR_LIB=gaustudio_b200/variants/libgsr_cluster_scan.so): with the default library GSR_SCAN_CLUSTER is ignored.
# failure rate of the sparse fused/un-fused test (in its test-file sequence, fresh process each) per switch setting;
# four processes at a time
mkdir -p gpurun_out/arms
K="not speculation and not tile_order and not graphed and not pipelined and not concurrent"
arm() {  # name cluster spec rounds
  for r in $(seq 1 $4); do
    for j in 1 2 3 4; do
      GSR_SCAN_CLUSTER=$2 GSR_SPECULATE=$3 timeout -k 5 90 python -m pytest tests/test_gpu_api.py -x -q -k "$K" > gpurun_out/arms/$1_${r}_$j.txt 2>&1 &
    done
    wait
  done
  f=$(grep -l "failed" gpurun_out/arms/$1_*.txt | wc -l); n=$(ls gpurun_out/arms/$1_*.txt | wc -l)
  echo "$1 (cluster=$2 speculate=$3): $f failed of $n"
  grep -h "E  .*AssertionError: fused.*re-render" gpurun_out/arms/$1_*.txt | cut -c1-260
  grep -h "^FAILED" gpurun_out/arms/$1_*.txt | sort | uniq -c
}
t0=$(date +%s)
arm off 0 0 4
arm on 1 1 3
arm cluster_only 1 0 3
arm spec_only 0 1 3
echo "total $(( $(date +%s) - t0 )) s"
