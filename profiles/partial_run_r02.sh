T=r02b; O=gpurun_out
echo "== e2e breakdown"; timeout 200 python profiles/e2e_resident.py 2>&1 | tee $O/${T}_e2e_resident.txt
echo "== memcheck(smoke)"; DRA_NO_PDL=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $O/${T}_memcheck.log 2>&1; echo "rc=$?"; tail -3 $O/${T}_memcheck.log
echo "== racecheck(smoke)"; DRA_NO_PDL=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > $O/${T}_racecheck.log 2>&1; echo "rc=$?"; tail -3 $O/${T}_racecheck.log
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches.csv python bench.py --steps 6 --warmup 3 --no-extras --no-resident > $O/${T}_ncu_bench.log 2>&1; grep -c "dra::" $O/${T}_launches.csv
echo "== ncu full: sort path, k_unsuitable, shard compaction"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bucket|k_pack|k_unsuitable|k_shard|k_fused" -s 4 -c 10 -f -o $O/prof_${T}_others python profiles/one_batch_sort.py > $O/${T}_ncu_others.log 2>&1; ls -la $O/prof_${T}_others.ncu-rep
