mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests/test_gpu_pdsch_enc.py -m gpu -x -q 2>&1 | tail -30
oracle/_ref/pdsch_hwacc_parity 200 2>&1 | tail -5
