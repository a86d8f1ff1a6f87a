#!/bin/bash
# quick GPU iteration: parity suite + headline bench (+ optional extra command in $1)
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --timeout 900 -x ) > gpurun_out/pytest.log 2>&1
tail -4 gpurun_out/pytest.log
python bench.py --steps 100 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
if [ -n "$1" ]; then bash -c "$1"; fi
