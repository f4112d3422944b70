set -u
mkdir -p gpurun_out/k2
python -m pytest tests -m gpu -q > gpurun_out/k2/tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/k2/tests.log
python __graft_entry__.py smoke 2>&1 | tail -1
{ for n in 256 512 1024 2048 4096; do python profiles/prof_sliding.py $n 2>&1 | grep "mode="; done; echo "# WAVESPEC_OVERLAP=0"; for n in 512 1024; do WAVESPEC_OVERLAP=0 python profiles/prof_sliding.py $n 2>&1 | grep "mode="; done; } > gpurun_out/k2/timings.txt; cat gpurun_out/k2/timings.txt
python bench.py > gpurun_out/k2/bench.json 2> gpurun_out/k2/bench.err; echo "bench rc=$?"; cat gpurun_out/k2/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k2/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/k2/ncu_launch.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sliding -c 1 -s 20 -o gpurun_out/k2/bench_kernel -f python bench.py --steps 2 --warmup 1 > gpurun_out/k2/ncu_full.log 2>&1; echo "ncu full rc=$?"
