mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_demod.py -m gpu -x -q --durations=5 > gpurun_out/pytest_demod.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_demod.log
