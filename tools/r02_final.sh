#!/bin/bash
# round-2 evidence run on one B200: full parity suite (config sizes included), headline bench lines (cosine, L2,
# reference arm), ncu launch list + full captures of the tensor path and the ADC scan, the other configs' bench lines,
# sanitizer passes.  Raw output under gpurun_out/r02final/ (scratch); tools/collect_profiles.py r02 copies the
# summaries into profiles/.
O=gpurun_out/r02final
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.txt
( time python -m pytest tests -m gpu -q --timeout 1500 ) > $O/pytest.log 2>&1
tail -3 $O/pytest.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --metric-kind l2 > $O/bench_n1_l2.json 2> $O/bench_n1_l2.err
python bench.py --impl reference --steps 3 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --workload c1 > $O/bench_c1.json 2>/dev/null
python bench.py --workload b1 > $O/bench_b1.json 2>/dev/null
python bench.py --workload c3 --steps 20 > $O/bench_c3.json 2> $O/bench_c3.err
python bench.py --workload c4 --steps 20 > $O/bench_c4.json 2> $O/bench_c4.err
python bench.py --workload c4 --steps 10 --batch 8192 > $O/bench_c4_b8192.json 2> $O/bench_c4_b8192.err
python bench.py --workload c5 --gpus 1 --rows 12500000 --metric-kind l2 --steps 10 > $O/bench_c5_one_shard.json 2> $O/bench_c5_one_shard.err
# launch list of the headline command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_tensor_path.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_tensor_path_l2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --metric-kind l2 > /dev/null 2>&1
# full captures: one step of the tensor path, the ADC scan
ncu --set full --clock-control none --import-source on -k regex:"flat_gemm_ts|ts_select|rescore|merge_topk|prep_queries" -s 24 -c 8 -o $O/prof_tensor python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_tensor.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"adc_ring|adc_scan" -s 2 -c 1 -o $O/prof_adc python tools/adc_sweep.py --n 2500000 --nlist 1024 --configs ";" > $O/ncu_adc.log 2>&1
for r in tensor adc; do python tools/ncu_summary.py $O/prof_$r.ncu-rep > $O/summary_$r.md 2>/dev/null; done
# sanitizers over every device path at small sizes
compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > $O/memcheck.log 2>&1; tail -3 $O/memcheck.log
compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_smoke.py > $O/racecheck.log 2>&1; tail -3 $O/racecheck.log
rm -f $O/prof_tensor.ncu-rep.tmp; ls -la $O
