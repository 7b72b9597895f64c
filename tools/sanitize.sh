#!/bin/bash
# compute-sanitizer over the GPU test-suite (the reference has no race / memory checker; this is the build's own).
#   memcheck  - out-of-bounds / misaligned global, shared and local accesses, leaks of device allocations
#   racecheck - shared-memory data races between the warps of a CTA (the hand-written kernels stage operands in shared memory)
#   synccheck - divergent / invalid barrier use
# Usage (on a GPU box):  tools/sanitize.sh [memcheck|racecheck|synccheck|all] [pytest -k expression]
# Writes gpurun_out/sanitize_<tool>.log; copy the summaries into profiles/ for the record. The tcgen05 / TMA kernels are included;
# the cooperative decode kernel runs under memcheck only (racecheck does not model cross-CTA polling and times out on it).
set -u
tool=${1:-memcheck}
expr=${2:-"not graphed and not baseline_sizes and not variant_b"}
mkdir -p gpurun_out
run() {
  local t=$1
  local extra=""
  # (no --leak-check: the stream-ordered caching allocator keeps its blocks until process exit by design; pdn_empty_cache releases them)
  [ "$t" != "memcheck" ] && expr="$expr and not llama and not plan and not reference_llama"
  echo "== compute-sanitizer --tool $t (pytest -m gpu -k \"$expr\")"
  PDN_SANITIZE=1 timeout 3000 compute-sanitizer --tool "$t" $extra --error-exitcode 9 --print-limit 20 \
    python -m pytest tests -m gpu -x -q -k "$expr" -p no:cacheprovider > gpurun_out/sanitize_$t.log 2>&1
  local rc=$?
  tail -n 12 gpurun_out/sanitize_$t.log
  grep -E "ERROR SUMMARY|LEAK SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitize_$t.log | tail -n 3
  echo "== $t rc=$rc"
  return $rc
}
if [ "$tool" = "all" ]; then
  rc=0
  for t in memcheck racecheck synccheck; do run $t || rc=$?; done
  exit $rc
fi
run "$tool"
