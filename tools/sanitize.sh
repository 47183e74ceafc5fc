#!/bin/bash
# compute-sanitizer over the hot path (the build-side race / memory checks): smoke() -- kernels 1-3
# incl. the tcgen05 / mbarrier / TMA pipeline and the lock-free run union-find -- under memcheck, racecheck, synccheck
# and initcheck.  Run on a GPU box:  gpurun --timeout 1500 -- 'bash tools/sanitize.sh'   -> gpurun_out/sanitizer_*.log
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
      python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke OK' gpurun_out/sanitizer_$tool.log | tr '\n' ' ')"
done
