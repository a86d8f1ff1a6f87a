#!/bin/bash
# how much of the GEMM's time is the epilogue?  dbg 1 = no accumulator scan, 2 = TMEM loads but no compare
for dbg in 0 1 2; do
  COMET_B200_DBG_EPI=$dbg ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_$dbg.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/l_$dbg.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h={n:i for i,n in enumerate(rows[hdr])}
g=[float(r[h["Metric Value"]])/1e3 for r in rows[hdr+1:] if len(r)>h["Metric Value"] and "flat_gemm" in r[h["Kernel Name"]]]
print("dbg=$dbg under ncu: gemm launches (us):", [round(x,1) for x in g[-3:]])
PY
  COMET_B200_DBG_EPI=$dbg python bench.py --steps 300 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['roofline']['step_share']; n=s['steps']
print('dbg=$dbg bench 300 steps: step %.3f ms gemm %.3f clocks %s' % (d['ms_per_step'], s['gemm_ms']/n, d['clocks']))"
done
