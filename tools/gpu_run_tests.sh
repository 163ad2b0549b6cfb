# Full GPU suite (parity tests proper, through the C ABI) + smoke.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 2700 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -18 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
