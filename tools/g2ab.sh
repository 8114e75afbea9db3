# A/B of the two-lane G2 round kernel against the one-lane kernel (variants/g2old.so) on one box; every step under its own timeout
for lib in collaborative-zksnark_b200/libczk_b200.so variants/g2old.so collaborative-zksnark_b200/libczk_b200.so; do
  echo -n "$lib: "; CZK_B200_LIB=$lib timeout 120 python tools/msm_once.py 2 20 0 2>&1 | grep -E "curve|rror" || echo "FAILED/timeout"
done
echo "== G2/MSM parity with the two-lane kernel"
timeout 300 python -m pytest tests/test_gpu_msm.py tests/test_gpu_large.py tests/test_golden_vectors.py -x -q -m gpu 2>&1 | tail -3
