#!/bin/bash
# compute-sanitizer passes over the small cases (run under gpurun on one B200; minutes, not seconds):
#   tools/sanitize.sh            memcheck + racecheck + initcheck of the smoke test and two fixture tests
# The reference has no race detection of its own; this is the B200 counterpart of its
# check_pairs() invariant: the force kernels never scatter, the list build writes only the atom's own rows.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  echo "== compute-sanitizer --tool $tool: smoke"
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python __graft_entry__.py smoke nobuild > gpurun_out/sanitize_${tool}_smoke.log 2>&1
  echo "   exit $?  ($(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_${tool}_smoke.log) summaries: $(grep 'ERROR SUMMARY' gpurun_out/sanitize_${tool}_smoke.log | tail -1))"
done
echo "== compute-sanitizer --tool memcheck: fixture tests (single species, two species, deformation)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x \
  -k "cu_nve-1 or nial_nvt-4 or cu_lindef-1 or cu_eeam-1 or two_species or npt_axial" > gpurun_out/sanitize_memcheck_fixtures.log 2>&1
echo "   exit $?  $(grep 'ERROR SUMMARY' gpurun_out/sanitize_memcheck_fixtures.log | tail -1)  $(tail -n 1 gpurun_out/sanitize_memcheck_fixtures.log)"
