#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sla_sanitize.py > gpurun_out/r2r_memcheck.log 2>&1; echo "memcheck exit $?"; tail -6 gpurun_out/r2r_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python tools/sla_sanitize.py > gpurun_out/r2r_racecheck.log 2>&1; echo "racecheck exit $?"; tail -12 gpurun_out/r2r_racecheck.log
