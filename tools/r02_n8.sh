#!/bin/bash
# 8-GPU run: BASELINE configs[4] (Flat L2, 100M x 768 = 8 x 12.5M rows, K=100, batch 512) as a 1/2/4/8 series through the
# library's single-process sharded index, then the driver's torchrun line at N=8 (replicas + rows_sharded sub-record)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n8.txt 2>&1
( time python bench.py --workload c5 --gpus 8 --rows 12500000 --metric-kind l2 --steps 20 ) > gpurun_out/r02_bench_c5_n8.json 2> gpurun_out/r02_bench_c5_n8.err
tail -6 gpurun_out/r02_bench_c5_n8.err; cat gpurun_out/r02_bench_c5_n8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
tail -3 gpurun_out/r02_bench_n8.err; cat gpurun_out/r02_bench_n8.json
( time python -m pytest tests/test_flat_sharded_gpu.py -m gpu -q --timeout 600 ) > gpurun_out/r02_pytest_sharded_n8.log 2>&1
tail -3 gpurun_out/r02_pytest_sharded_n8.log
