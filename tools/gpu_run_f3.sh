mkdir -p gpurun_out
oracle/_ref/upper_phy_wiring | tee gpurun_out/r2_upper_phy_wiring.txt
timeout 1200 python -m pytest tests/test_gpu_integration.py -m gpu -x -q 2>&1 | tail -15
