set -u
export PYTHONPATH=.
timeout 420 python -m pytest tests/test_gpu_multi.py -q --tb=short > gpurun_out/n2d_pytest.log 2>&1; tail -12 gpurun_out/n2d_pytest.log
