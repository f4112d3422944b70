set -u
python -m pytest tests -m gpu -x -q > gpurun_out/k1_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/k1_tests.log
python __graft_entry__.py smoke > gpurun_out/k1_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/k1_smoke.log
for n in 512 1024 2048; do
 for k in "" "128,2" "128,4" "256,2" "256,4" "256,8" "512,8" "512,16"; do
  echo -n "K1=$k  "; WAVESPEC_K1=$k python profiles/prof_k1.py $n 2>&1 | tail -1
 done
done
