#!/bin/bash
# sweep of the tensor path's phase plan (rows of phase A x growth factor): queries/s of the headline step
for ra in 1024 2048 3072 5000 8192; do
  for gr in 6 10 16 24; do
    v=$(COMET_B200_ROWS_A=$ra COMET_B200_PHASE_GROWTH=$gr python bench.py --steps 60 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['roofline']['step_share']; n=s['steps']
print('%.0f q/s  step %.3f ms  gemm %.3f sel %.3f resc %.3f passes %d' % (d['value'], d['ms_per_step'], s['gemm_ms']/n, s['select_ms']/n, s['rescore_ms']/n, d['config']['scan_passes_per_step']))")
    echo "rows_a=$ra growth=$gr : $v"
  done
done
