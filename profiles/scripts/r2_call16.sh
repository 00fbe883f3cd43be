#!/bin/bash
# Round-2 GPU call 16: compute-sanitizer (memcheck, then racecheck on the packed cases) over every kernel incl. the
# packed-rows layout, the item-stream attention backward and the several-rows-per-warp row kernels.
mkdir -p gpurun_out/r16
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python tests/sanitize_smoke.py > gpurun_out/r16/memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|^ok|Error|Invalid" gpurun_out/r16/memcheck.log | head -30
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "
import sys; sys.path.insert(0, '.')
import tests.sanitize_smoke as t
t.packed_rows()" > gpurun_out/r16/racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|^ok|hazard" gpurun_out/r16/racecheck.log | head -20
