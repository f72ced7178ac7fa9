#!/usr/bin/env bash
# compute-sanitizer over the torch-free self test, ONE tool per gpurun call (B200_PROFILING.md: running several tools in
# one call has left GPUs unusable).  Usage on a B200 box, only after `tools/cabi_selftest` itself has exited 0 there:
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh memcheck'      (or: racecheck | synccheck | initcheck)
# Small shapes only (the sanitizer slows kernels down by 10-100x): the shape sweep + the host entry + the exchange phases.
# NOTE (round 2): compute-sanitizer is closed on this pool (it answers with a notice and runs nothing); the script is kept for
# boxes where it is available.  Races were reviewed by hand instead (barrier placement is commented at every shared-memory hand-off).
set -u
tool=${1:-memcheck}
mkdir -p gpurun_out
SELFTEST_ONLY=sweep,host,exchange timeout 300 tools/cabi_selftest 20000 20000 > gpurun_out/sanitize_plain.log 2>&1 &&
SELFTEST_ONLY=sweep,host,exchange timeout 800 compute-sanitizer --tool "$tool" --error-exitcode 9 \
  tools/cabi_selftest 20000 20000 > "gpurun_out/sanitize_$tool.log" 2>&1
echo "rc=$?"
grep -E "ERROR SUMMARY|Invalid|Race|hazard|MISMATCH|selftest" "gpurun_out/sanitize_$tool.log" | tail -20
