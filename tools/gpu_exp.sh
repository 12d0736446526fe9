#!/bin/bash
for e in 1 3 7; do echo "=== EVK_TC_EXP=$e"; EVK_TC_EXP=$e EVK_TC_TIMING=1 FRAMES=3 BATCH=24 timeout 120 python tools/tc_experiment.py 2>&1 | grep TIMING | sort -u -k1,12 | cut -c1-210; done
