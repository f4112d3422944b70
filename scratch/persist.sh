timeout 120 python profiles/prof_sliding.py 2>&1 | grep "mode=both"
WAVESPEC_PERSIST=0 timeout 120 python profiles/prof_sliding.py 2>&1 | grep "mode=both"
timeout 120 python profiles/prof_sliding.py 512 2>&1 | grep "mode=both"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sliding or config2 or full_size" 2>&1 | tail -2
