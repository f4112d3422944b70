mkdir -p gpurun_out/k2
SEC="--section WarpStateStats --section SourceCounters --section SchedulerStats --section LaunchStats --section Occupancy --section SpeedOfLight"
ncu $SEC --clock-control none --import-source on -k regex:sliding_persistent -c 1 -s 1 -o gpurun_out/k2/persist -f python profiles/prof_sliding.py > gpurun_out/k2/ncu_p.log 2>&1; echo "rc=$?"
WAVESPEC_PERSIST=0 ncu $SEC --clock-control none --import-source on -k regex:sliding_overlap -c 1 -s 1 -o gpurun_out/k2/overlap -f python profiles/prof_sliding.py > gpurun_out/k2/ncu_o.log 2>&1; echo "rc=$?"
ls -la gpurun_out/k2/
