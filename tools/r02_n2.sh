#!/bin/bash
# 2-GPU iteration: sharded tests over real peers, the torchrun headline line with its rows_sharded sub-record, the c5 series
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
( time python -m pytest tests/test_flat_sharded_gpu.py -m gpu -q --timeout 600 ) > gpurun_out/r02_pytest_sharded_n2.log 2>&1
tail -5 gpurun_out/r02_pytest_sharded_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -3 gpurun_out/r02_bench_n2.err; cat gpurun_out/r02_bench_n2.json
( time python bench.py --workload c5 --gpus 2 --rows ${C5_ROWS:-12500000} --metric-kind l2 --steps 10 ) > gpurun_out/r02_bench_c5_n2.json 2> gpurun_out/r02_bench_c5_n2.err
tail -5 gpurun_out/r02_bench_c5_n2.err; cat gpurun_out/r02_bench_c5_n2.json
