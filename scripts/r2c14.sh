set -u
export PYTHONPATH=.
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "model_zoo or stem_tensorcore or dropout" 2>&1 | tail -25
