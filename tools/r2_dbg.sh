#!/bin/bash
mkdir -p gpurun_out
CPHB_DEBUG_EVENTS=1 CPHB_DEBUG_CERT=1 timeout 120 python tools/one_registration.py --warm 1 > gpurun_out/dbg_a.txt 2>&1; grep -A1 "timeline\|per launch" gpurun_out/dbg_a.txt | cut -c1-700
