#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool initcheck --print-limit 40 --error-exitcode 9 python __graft_entry__.py smoke nobuild > gpurun_out/sanitize_initcheck_smoke.log 2>&1
echo "exit $?"; grep "ERROR SUMMARY" gpurun_out/sanitize_initcheck_smoke.log | tail -2
grep -A14 "Uninitialized" gpurun_out/sanitize_initcheck_smoke.log | grep -E "Uninitialized|at |Host Frame: imdb|Device Frame" | head -40 | cut -c1-220
