#!/bin/bash
# round-end evidence run on one B200: parity suite, bench lines, ncu launch lists and full captures.
# Writes raw output under gpurun_out/ (scratch); tools/collect_profiles.py turns it into profiles/.
O=gpurun_out/final
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.txt
( time python -m pytest tests -m gpu -q --timeout 900 ) > $O/pytest.log 2>&1
tail -3 $O/pytest.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --path exact --steps 3 --no-cpu-baseline > $O/bench_exact.json 2>/dev/null
python bench.py --workload c1 > $O/bench_c1.json 2>/dev/null
python bench.py --workload b1 > $O/bench_b1.json 2>/dev/null
python bench.py --rows 12500000 --metric-kind l2 --steps 10 --no-cpu-baseline > $O/bench_shard_12.5M_l2.json 2>/dev/null
# launch lists (same commands as the bench, short)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_tensor_path.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_exact_scan.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --path exact > /dev/null 2>&1
# full captures
ncu --set full --clock-control none --import-source on -k regex:"flat_gemm|cand_select|rescore|merge_topk" -s 28 -c 8 -o $O/prof_tensor python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_tensor.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:flat_scan -s 100 -c 1 -o $O/prof_scan python bench.py --steps 2 --warmup 3 --no-cpu-baseline --path exact > $O/ncu_scan.log 2>&1
python tools/bench_indexes.py --n 1000000 --only ivf,pq,ivfpq > $O/idx_1m.json 2> $O/idx_1m.err
python tools/bench_indexes.py --only c3 > $O/c3.json 2> $O/c3.err
python tools/bench_indexes.py --only c4 > $O/c4.json 2> $O/c4.err
ncu --set full --clock-control none --import-source on -k regex:"adc_scan" -s 2 -c 1 -o $O/prof_adc python tools/bench_indexes.py --only c3 --c3-n 2500000 --c3-nlist 1024 > $O/ncu_adc.log 2>&1
for r in tensor scan adc; do python tools/ncu_summary.py $O/prof_$r.ncu-rep > $O/summary_$r.md 2>/dev/null; done
ls -la $O
