timeout 300 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu 2>&1 | tail -2
timeout 120 python tools/ntt_once.py 21 2>&1 | tail -12
