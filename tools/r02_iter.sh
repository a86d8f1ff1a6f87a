#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_flat_tensor_gpu.py -m gpu -q -x --timeout 900 ) > gpurun_out/r02_pytest_tensor.log 2>&1
tail -3 gpurun_out/r02_pytest_tensor.log
timeout 300 python tools/dbg_gemm.py A=0 A=1 > gpurun_out/r02_probe2.txt 2>&1
DBG_METRIC=l2 timeout 300 python tools/dbg_gemm.py A=0 COMET_B200_DBG_EPI=1 >> gpurun_out/r02_probe2.txt 2>&1
cat gpurun_out/r02_probe2.txt
M=gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum
DBG_METRIC=l2 timeout 300 ncu --metrics $M --clock-control none -k regex:flat_gemm -s 12 -c 3 --csv --log-file gpurun_out/r02_ncu_ts_l2.csv python tools/dbg_gemm.py A=0 > /dev/null 2>&1
