set -u
python profiles/prof_k1.py 1024 | tail -1
ncu --set full --clock-control none --import-source on -k regex:window_fft -c 1 -s 2 -o gpurun_out/k1w_1024 -f python profiles/prof_k1.py 1024 > gpurun_out/k1_ncu.log 2>&1; echo "ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:window_fft -c 1 -s 2 -o gpurun_out/k1w_2048 -f python profiles/prof_k1.py 2048 > gpurun_out/k1_ncu2.log 2>&1; echo "ncu rc=$?"
