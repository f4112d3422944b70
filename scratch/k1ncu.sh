set -u
python -m pytest tests -m gpu -x -q > gpurun_out/k1_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/k1_tests.log
python profiles/prof_k1.py 1024 | tail -1
python profiles/prof_k1.py 512 | tail -1
ncu --set full --clock-control none --import-source on -k regex:window_fft -c 1 -s 2 -o gpurun_out/k1_1024 -f python profiles/prof_k1.py 1024 > gpurun_out/k1_ncu.log 2>&1; echo "ncu rc=$?"
