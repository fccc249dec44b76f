#!/bin/bash
# compute-sanitizer over the op-level kernels (run under gpurun, ONE GPU): racecheck (shared-memory hazards of the
# hand-rolled mbarrier / named-barrier protocols), synccheck (barrier misuse), memcheck (out-of-bounds).
# Each tool runs the bring-up groups in one process; logs land in gpurun_out/sanitizer_<tool>.log.
GROUPS_TO_RUN=${GROUPS_TO_RUN:-"gemm_basic gemm_epi conv attention misc"}
for tool in ${TOOLS:-racecheck synccheck memcheck}; do
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 9 \
    python tools/bringup.py $GROUPS_TO_RUN > gpurun_out/sanitizer_$tool.log 2>&1
  echo "sanitizer $tool exit=$?" | tee -a gpurun_out/sanitizer_$tool.log
  tail -5 gpurun_out/sanitizer_$tool.log
done
