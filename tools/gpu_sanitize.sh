#!/bin/bash
# compute-sanitizer over the kernels added in the second half of round 1 (wide path, TEA pair kernel, analysis)
mkdir -p gpurun_out
export OMP_NUM_THREADS=4
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -q -m gpu -x -k "wide_path_equals_cta_path_bitwise and mt40 or tea_window or device_analysis_on_broken or wide_path_reports" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/memcheck.log | tail -8
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests -q -m gpu -x -k "tea_window or device_analysis_on_broken" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/racecheck.log | tail -8
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 3 python -m pytest tests -q -m gpu -x -k "wide_large_tea" > gpurun_out/initcheck.log 2>&1; echo "initcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Uninitialized" gpurun_out/initcheck.log | tail -8
