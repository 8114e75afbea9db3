# lanes + device-side rounds: correctness first, then timings; every step under its own timeout
timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_gpu_groth16.py tests/test_gpu_large.py tests/test_golden_vectors.py tests/test_gpu_gsz.py -x -q -m gpu 2>&1 | tail -3
for w in 24 12 6 3; do
  echo -n "G2 walk=$w: "; CZK_BAT_WALK_G2=$w timeout 120 python tools/msm_once.py 2 20 0 2>&1 | grep -E "curve|rror" || echo FAILED
done
for w in 24 12 6; do
  echo -n "G1 walk=$w: "; CZK_BAT_WALK_G1=$w timeout 120 python tools/msm_once.py 1 21 0 2>&1 | grep -E "curve|rror" || echo FAILED
done
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_lanes.json 2> gpurun_out/r2_bench_lanes.err; tail -c 1500 gpurun_out/r2_bench_lanes.json; tail -3 gpurun_out/r2_bench_lanes.err
