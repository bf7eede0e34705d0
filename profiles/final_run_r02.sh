#!/bin/bash
# All single-GPU artefacts of round 2 in ONE gpurun call:  bash profiles/final_run_r02.sh r02
# Outputs go to gpurun_out/<tag>_*; profiles/collect_r02.sh copies what should be judged into profiles/.
T=${1:-r02}; O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/${T}_pytest_gpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py 2>$O/${T}_bench.err | tail -1 > $O/${T}_bench.json; cut -c1-300 $O/${T}_bench.json
timeout 300 python bench.py --no-resident --no-extras 2>/dev/null | tail -1 > $O/${T}_bench_coop_launch.json
timeout 300 python bench.py --no-direct --no-extras 2>/dev/null | tail -1 > $O/${T}_bench_copyengine.json
timeout 300 python bench.py --impl reference 2>/dev/null | tail -1 > $O/${T}_bench_reference.json; cut -c1-200 $O/${T}_bench_reference.json
echo "== e2e breakdown"; timeout 200 python profiles/e2e_resident.py 2>&1 | tee $O/${T}_e2e_resident.txt
echo "== timeline"; DRA_TIMELINE=1 timeout 200 python profiles/timeline.py > $O/timeline_${T}.txt 2>&1; tail -13 $O/timeline_${T}.txt
# (compute-sanitizer does not honour programmatic dependent launches: a dependent kernel may run before its producer has
#  finished, which it never does outside the tool — the sort path is therefore checked with DRA_NO_PDL=1)
echo "== memcheck(smoke)"; DRA_NO_PDL=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $O/${T}_memcheck.log 2>&1; echo "rc=$?"; tail -3 $O/${T}_memcheck.log
echo "== racecheck(smoke)"; DRA_NO_PDL=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > $O/${T}_racecheck.log 2>&1; echo "rc=$?"; tail -3 $O/${T}_racecheck.log
echo "== memcheck(smoke) with PDL, for the record"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $O/${T}_memcheck_pdl.log 2>&1; echo "rc=$?"; tail -2 $O/${T}_memcheck_pdl.log
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches.csv python bench.py --steps 6 --warmup 3 --no-extras --no-resident > $O/${T}_ncu_bench.log 2>&1; grep -c "dra::" $O/${T}_launches.csv
echo "== ncu full: k_fused"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -f -o $O/prof_${T}_fused python profiles/one_batch.py > $O/${T}_ncu_full.log 2>&1; ls -la $O/prof_${T}_fused.ncu-rep
echo "== ncu full: sort path, k_unsuitable, shard compaction"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bucket|k_pack|k_unsuitable|k_shard|k_fused" -s 4 -c 10 -f -o $O/prof_${T}_others python profiles/one_batch_sort.py > $O/${T}_ncu_others.log 2>&1; ls -la $O/prof_${T}_others.ncu-rep
echo "== sort-path stage times (cfg3, cfg5, 1M)"; timeout 300 python profiles/stage_times.py 2>&1 | tail -3 | tee $O/${T}_stage_times.txt
echo "== ncu full: large batch (1M claims)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bucket|k_pack" -s 8 -c 4 -f -o $O/prof_${T}_large python profiles/one_batch_large.py > $O/${T}_ncu_large.log 2>&1; ls -la $O/prof_${T}_large.ncu-rep
