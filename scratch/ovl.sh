python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for t in "32,2" "32,1" "32,4" "28,1" "24,1" "24,2"; do echo "TILE=$t"; WAVESPEC_TILE=$t python profiles/prof_sliding.py 2>&1 | grep "mode=both"; done
for t in "64,4" "64,2" "32,2" "32,1" "48,2" "48,1"; do echo "N=512 TILE=$t"; WAVESPEC_TILE=$t python profiles/prof_sliding.py 512 2>&1 | grep "mode=both"; done
