for t in "32,4" "32,2" "32,1" "32,4" "32,2" "32,1"; do echo "TILE=$t"; WAVESPEC_TILE=$t python profiles/prof_sliding.py 2>&1 | grep "mode="; done
