#!/bin/bash
for a in "256 256 128 up" "128 256 128 up" "64 512 256 up" "128 32 256 up" "512 128 128 conv3"; do echo "== $a"; timeout 120 python tools/tc_timing.py $a 2>&1 | tail -7; done
