#!/bin/bash
mkdir -p gpurun_out
python tools/dbg_cos.py > gpurun_out/r02_dbg_cos.txt 2>&1
grep -c "fallback 0" gpurun_out/r02_dbg_cos.txt; grep -v "fallback 0\|staged" gpurun_out/r02_dbg_cos.txt
( timeout 1200 python -m pytest tests/test_flat_tensor_gpu.py tests/test_flat_gpu.py -m gpu -q -x --timeout 900 ) > gpurun_out/r02_pytest_tensor.log 2>&1
tail -5 gpurun_out/r02_pytest_tensor.log
timeout 300 python tools/dbg_gemm.py COMET_B200_TAIL=0 COMET_B200_TAIL=0 > gpurun_out/r02_probe2.txt 2>&1
cat gpurun_out/r02_probe2.txt
