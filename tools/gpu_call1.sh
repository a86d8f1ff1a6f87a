#!/bin/bash
# one GPU-box session: parity suite, headline bench + reference arm, select-kernel source profile, ADC numbers
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
( time python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/pytest.log 2>&1
tail -5 gpurun_out/pytest.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --path exact --steps 3 --no-cpu-baseline > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cand_select" -s 9 -c 3 -o gpurun_out/prof_select \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_select.log 2>&1
python tools/bench_indexes.py --n 1000000 --only pq,ivfpq > gpurun_out/idx_1m.json 2> gpurun_out/idx_1m.err
timeout 900 python tools/bench_indexes.py --only c3 > gpurun_out/c3.json 2> gpurun_out/c3.err
timeout 600 python tools/bench_indexes.py --only c4 > gpurun_out/c4.json 2> gpurun_out/c4.err
cat gpurun_out/idx_1m.json gpurun_out/c3.json gpurun_out/c4.json
