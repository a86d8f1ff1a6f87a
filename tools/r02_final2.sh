#!/bin/bash
# round-2 closing evidence on one B200 with the final build: the whole parity suite, smoke, the headline line and the
# reference arm, the IVFPQ / HNSW / 10K config lines, racecheck over every device path.  Raw output under
# gpurun_out/r02final2/ (scratch); the summaries are copied into profiles/ by hand (names r02_final2_*).
O=gpurun_out/r02final2
mkdir -p $O
( time python -m pytest tests -m gpu -q --timeout 1500 ) > $O/pytest.log 2>&1
tail -3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference --steps 3 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --workload c1 > $O/bench_c1.json 2>/dev/null
python bench.py --workload c3 --steps 20 > $O/bench_c3.json 2> $O/bench_c3.err
python bench.py --workload c4 --steps 20 > $O/bench_c4.json 2> $O/bench_c4.err
for f in bench_n1 bench_c1 bench_c3 bench_c4; do python - <<PY
import json
d=json.load(open("$O/$f.json"))
print("$f", round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"],4), "parity", d.get("parity"))
PY
done
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_smoke.py > $O/racecheck.log 2>&1; tail -3 $O/racecheck.log
