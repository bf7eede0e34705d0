#!/bin/bash
# All single-GPU artefacts of a round in ONE gpurun call:  bash profiles/final_run.sh r01f
# Outputs go to gpurun_out/<tag>_*; copy what should be judged into profiles/.
T=${1:-rXX}; O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== memcheck(smoke)"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $O/${T}_memcheck.log 2>&1; echo "rc=$?"; tail -3 $O/${T}_memcheck.log
echo "== bench"; python bench.py 2>$O/${T}_bench.err | tail -1 > $O/${T}_bench.json; cat $O/${T}_bench.json | cut -c1-400
python bench.py --no-direct 2>/dev/null | tail -1 > $O/${T}_bench_copyengine.json
python bench.py --impl reference 2>/dev/null | tail -1 > $O/${T}_bench_reference.json; cut -c1-300 $O/${T}_bench_reference.json
echo "== configs"; python bench_configs.py > $O/${T}_configs.jsonl 2>$O/${T}_configs.err; python -c "
import json
for l in open('$O/${T}_configs.jsonl'):
    d=json.loads(l); print({k:d[k] for k in list(d)[:9]})"
echo "== unsuitable"; python bench_unsuitable.py > $O/${T}_unsuitable.json 2>/dev/null; cut -c1-400 $O/${T}_unsuitable.json
echo "== timeline"; DRA_TIMELINE=1 python profiles/timeline.py > $O/timeline_${T}.txt 2>&1; tail -13 $O/timeline_${T}.txt
echo "== ncu launch list"; ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python bench.py --steps 6 --warmup 3 > $O/${T}_ncu_bench.log 2>&1; grep -c "dra::" $O/${T}_launches.csv
echo "== ncu full capture of k_fused"; ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -f -o $O/prof_${T}_fused python profiles/one_batch.py > $O/${T}_ncu_full.log 2>&1; ls -la $O/prof_${T}_fused.ncu-rep
