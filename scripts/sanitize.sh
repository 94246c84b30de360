#!/bin/bash
# compute-sanitizer passes over small GPU tests (memcheck, racecheck on shared memory, synccheck)
set -u
mkdir -p gpurun_out
K='fault_only or viscoelastic_machinery or dilatancy or zero_copy or vcabm5_decay or decay_problem or non_power_of_two'
for tool in memcheck racecheck synccheck; do
  echo "== $tool" | tee -a gpurun_out/sanitizer.log
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_rhs.py tests/test_gpu_solve.py -m gpu -q -x -k "$K" 2>&1 \
    | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -20 | tee -a gpurun_out/sanitizer.log
done
