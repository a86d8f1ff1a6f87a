#!/bin/bash
# 8-GPU run, second part: IVF / IVFPQ list shards against the single-GPU index, the sharded parity tests over real peers,
# the driver's torchrun line at N = 2 / 4 / 8 with the final build, configs[4] once more
mkdir -p gpurun_out
( time python -m pytest tests/test_flat_sharded_gpu.py tests/test_ivf_sharded_gpu.py -m gpu -q --timeout 900 ) > gpurun_out/r02_pytest_sharded_n8b.log 2>&1
tail -3 gpurun_out/r02_pytest_sharded_n8b.log
python tools/list_shards_bench.py --n 2000000 > gpurun_out/r02_list_shards_n8.json 2> gpurun_out/r02_list_shards_n8.err
tail -2 gpurun_out/r02_list_shards_n8.err | cut -c1-600
for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 3 > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err
python -c "import json; d=json.load(open('gpurun_out/r02_scale_n$n.json')); r=d['rows_sharded']; print($n, round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), 'rows_sharded', round(r['ms_per_step'],4), round(r['merge_ms'],4), round(r['slowest_shard_search_ms'],4), round(r['e2e']['ms_per_step'],3))"
done
python bench.py --workload c5 --gpus 8 --rows 12500000 --metric-kind l2 --steps 20 > gpurun_out/r02_bench_c5_n8b.json 2> gpurun_out/r02_bench_c5_n8b.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_c5_n8b.json'))
for r in d['series']: print(r['n_gpus'], round(r['ms_per_step'],3), round(r['value']), round(r['weak_scaling_efficiency_vs_1gpu'],3), round(r['merge_ms'],4), round(r['e2e']['ms_per_step'],3))
"
