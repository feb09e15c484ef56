set -u
export PYTHONPATH=.
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "follow_server or k_of_p or timeout or kill_mode or push_reduce or weights_ready or model_zoo" 2>&1 | tail -15
