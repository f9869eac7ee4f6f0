#!/bin/bash
# first run of the mirrored-tile forward: accuracy / timing check, then the histogram tests and the bench line
mkdir -p gpurun_out
timeout 300 python tools/tc_check_sym.py > gpurun_out/sym_check.log 2>&1; echo "sym check rc=$?"; cat gpurun_out/sym_check.log | tail -40
