mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_parity_ref.py -m gpu -q --durations=8 > gpurun_out/pytest_ref.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_ref.log
