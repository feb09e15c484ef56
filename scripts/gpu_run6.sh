#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tc_" > gpurun_out/pytest_gpu6.log 2>&1; grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu6.log | tail -12 | cut -c1-200
timeout 600 python scripts/bench_conv.py > gpurun_out/bench_conv.log 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/bench_conv.log'):
    if l.startswith('{'):
        r=json.loads(l); print(r['shape'], 'cudnn %.1f tap %.1f halo %.1f halo2 %.1f us | err tap %s halo %s halo2 %s %s' % (r['cudnn_us'], r.get('tap_us',-1), r.get('halo_us',-1), r.get('halo2_us',-1), r.get('tap_maxerr'), r.get('halo_maxerr'), r.get('halo2_maxerr'), r.get('halo2_err','')), '| wgrad cudnn %.1f ours %.1f err %s %s' % (r.get('cudnn_wgrad_us',-1), r.get('wgrad_us',-1), r.get('wgrad_maxerr'), r.get('wgrad_err','')))
PY
grep -E '^E  |Error' gpurun_out/pytest_gpu6.log | head -12 | cut -c1-250
