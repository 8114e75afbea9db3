timeout 300 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu 2>&1 | tail -2
echo "== tile 2^11, 256 threads, 2 blocks/SM"; timeout 120 python tools/ntt_once.py 21 2>&1 | grep -E "x6|fft  *x1"
echo "== tile 2^10, 128 threads, 4 blocks/SM"; CZK_B200_LIB=variants/ntt10.so timeout 120 python tools/ntt_once.py 21 2>&1 | grep -E "x6|fft  *x1"
CZK_B200_LIB=variants/ntt10.so timeout 300 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu 2>&1 | tail -2
