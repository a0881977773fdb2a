#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2g_parity.log
for mode in 0 1 2; do DFD_LN_FUSE=$mode timeout 300 python tools/parity_report.py 2>&1 | grep DFD_LN_FUSE >> gpurun_out/r2g_parity.log; done
cat gpurun_out/r2g_parity.log
