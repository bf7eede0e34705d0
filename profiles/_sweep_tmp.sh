timeout 200 python profiles/stage_times.py 2>&1 | tail -3
for t in 4096 16384 32768; do echo "T=$t"; DRA_HIST_TILE=$t timeout 200 python profiles/stage_times.py 2>&1 | tail -1; done
