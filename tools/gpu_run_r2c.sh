mkdir -p gpurun_out
bash tools/ab_bench.sh > gpurun_out/r2c_ab.txt 2>&1; cat gpurun_out/r2c_ab.txt
timeout 1500 python -m pytest tests/test_gpu_parity_ref.py -m gpu -x -q --durations=10 > gpurun_out/r2c_pytest_ref.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2c_pytest_ref.log
