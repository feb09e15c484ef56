#!/bin/bash
# compute-sanitizer over one complete training step (the reference has no race detection at all).  `smoke()` runs six steps
# of ResNet-18 (batch 64) through the engine: every tcgen05 kernel (stem, CTA-pair halo conv, tap-table conv incl. strided dgrad, the three wgrad
# kernels, classifier GEMMs), the BN / pooling / loss kernels and the PS kernels (push, reduce-update, broadcast, wait-flag) run under the tool.
# racecheck = shared-memory hazards, synccheck = barrier misuse, memcheck = out-of-bounds / misaligned; logs -> gpurun_out/sanitize_*.log
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
export PYTHONPATH=.
for tool in ${TOOLS:-synccheck racecheck memcheck}; do
  timeout ${SAN_TIMEOUT:-420} $CS --tool $tool --error-exitcode 7 --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_$tool.log
  grep -hE "ERROR SUMMARY|smoke ok|Error|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -12
done
