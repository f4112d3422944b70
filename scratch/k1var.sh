set -u
python -m pytest tests -m gpu -x -q > gpurun_out/k1_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/k1_tests.log
for v in 1 2; do for n in 512 1024 2048; do echo -n "K1W=$v "; WAVESPEC_K1W=$v python profiles/prof_k1.py $n | tail -1; done; done
