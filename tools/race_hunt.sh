#!/bin/bash
for cl in 1 0; do
  GSR_SCAN_CLUSTER=$cl timeout -k 5 100 python tools/race_hunt.py 40000 40 2>&1 | grep -v Warning | tail -12
  GSR_SCAN_CLUSTER=$cl timeout -k 5 100 python tools/race_hunt.py 150000 30 2>&1 | grep -v Warning | tail -12
done
