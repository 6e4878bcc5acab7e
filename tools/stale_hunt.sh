#!/bin/bash
GSR_PRIORITY=1 timeout -k 5 100 python tools/stale_hunt.py 120 2>&1 | grep -v Warning | tail -12
GSR_PRIORITY=0 timeout -k 5 100 python tools/stale_hunt.py 120 2>&1 | grep -v Warning | tail -12
