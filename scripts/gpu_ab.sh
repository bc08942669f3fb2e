#!/bin/bash
set -u
mkdir -p gpurun_out
MX_SCAN_TC_PROF=1 python scripts/scan_tc_prof.py > gpurun_out/scan_tc_prof.txt 2>&1; cat gpurun_out/scan_tc_prof.txt | tail -20
