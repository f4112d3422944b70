set -u
mkdir -p gpurun_out/k1
python -m pytest tests -m gpu -q > gpurun_out/k1/tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/k1/tests.log
{
echo "# per-window FFT path: IIR detrend + Blackman, spectra + bins, 4 series x 60k bars (prof_k1.py)"
for n in 256 512 1024 2048 4096; do python profiles/prof_k1.py $n | tail -1; done
echo "# same with WAVESPEC_K1W=0 (CTA-per-window-group kernel only)"
for n in 512 1024 2048; do WAVESPEC_K1W=0 python profiles/prof_k1.py $n | tail -1; done
echo "# BASELINE config shapes (prof_configs.py)"
python profiles/prof_configs.py 2>&1 | tail -22
} > gpurun_out/k1/timings.txt 2>&1
cat gpurun_out/k1/timings.txt
for n in 512 1024 2048; do
ncu --set full --clock-control none --import-source on -k regex:window_fft -c 1 -s 2 -o gpurun_out/k1/k1w_$n -f python profiles/prof_k1.py $n > gpurun_out/k1/ncu_$n.log 2>&1; echo "ncu $n rc=$?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/k1/launches_k1.csv python profiles/prof_k1.py 2048 > /dev/null 2>&1; echo "launch list rc=$?"
