set -u
python -m pytest tests -m gpu -x -q > gpurun_out/k1_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/k1_tests.log
for n in 256 512 1024 2048 4096; do python profiles/prof_k1.py $n | tail -1; done
python profiles/prof_configs.py 2>&1 | tail -12
