set -u
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_multi.py -q -k "sharded_and_ddp or resume_on_nvlink" 2>&1 | tail -8
timeout 600 python scripts/overlap_noise.py 32 2>&1 | tail -12
