#!/usr/bin/env bash
# Local helper (runs HERE, not on the GPU box): keep asking gpurun for a slot until the call is actually run.
#   bash scripts/gpu_retry.sh <out-file> [gpurun args...] -- '<command>'
# gpurun exits 3 / prints status=transient when the pod has no free slot (nothing charged): wait and retry, up to TRIES times.
out=$1; shift
for i in $(seq 1 ${TRIES:-40}); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  if ! grep -q "status=transient" "$out" && [ $rc -ne 3 ]; then echo "done rc=$rc try=$i" >> "$out"; exit $rc; fi
  sleep ${WAIT:-90}
done
echo "gave up after $i tries" >> "$out"
