#!/bin/bash
TAG=${1:-san}
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_run.py > gpurun_out/${TAG}_memcheck.txt 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/${TAG}_memcheck.txt
grep -c "Invalid\|ERROR SUMMARY" gpurun_out/${TAG}_memcheck.txt; tail -12 gpurun_out/${TAG}_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_run.py > gpurun_out/${TAG}_racecheck.txt 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/${TAG}_racecheck.txt
tail -6 gpurun_out/${TAG}_racecheck.txt
