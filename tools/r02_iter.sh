#!/bin/bash
mkdir -p gpurun_out
free -g | head -2; nproc
( time timeout 2400 python -m pytest tests/test_config_sizes_gpu.py -m gpu -q -x --timeout 1500 --durations=8 ) > gpurun_out/r02_pytest_sizes.log 2>&1
tail -25 gpurun_out/r02_pytest_sizes.log
