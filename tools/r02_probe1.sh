#!/bin/bash
# r02 probe 1: where does the tensor-path GEMM time go?  (a) baseline, (b) no corpus loads, (c) no query loads,
# (d) no loads at all, (e) epilogue off; plus cuBLAS on the same shape.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r02_probe1.txt
python tools/dbg_gemm.py COMET_B200_DBG_EPI=0 COMET_B200_DBG_EPI=64 COMET_B200_DBG_EPI=128 COMET_B200_DBG_EPI=192 COMET_B200_DBG_EPI=1 COMET_B200_DBG_EPI=193 COMET_B200_DBG_EPI=0 >> gpurun_out/r02_probe1.txt 2>&1
python tools/cublas_shape.py >> gpurun_out/r02_probe1.txt 2>&1
cat gpurun_out/r02_probe1.txt
