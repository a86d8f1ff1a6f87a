#!/bin/bash
# r02 iteration: tensor-path parity tests, then the GEMM probe
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_flat_tensor_gpu.py -m gpu -q -x --timeout 600 ) > gpurun_out/r02_pytest_tensor.log 2>&1
tail -5 gpurun_out/r02_pytest_tensor.log
timeout 300 python tools/dbg_gemm.py COMET_B200_DBG_EPI=0 COMET_B200_DBG_EPI=1 COMET_B200_DBG_EPI=65 COMET_B200_DBG_EPI=0 > gpurun_out/r02_probe2.txt 2>&1
cat gpurun_out/r02_probe2.txt
