#!/bin/bash
echo "== auto"; timeout 200 python tools/prof_blur.py 4 2>&1 | tail -13
echo "== big=0"; HFAGP_BLUR_BIG=0 timeout 200 python tools/prof_blur.py 4 2>&1 | tail -13
echo "== big=1"; HFAGP_BLUR_BIG=1 timeout 200 python tools/prof_blur.py 4 2>&1 | tail -13
