timeout 500 python -m pytest tests/test_gpu_msm.py tests/test_golden_vectors.py tests/test_gpu_large.py tests/test_gpu_setup.py -x -q -m gpu 2>&1 | tail -2
for args in "1 21" "1 20" "1 18" "2 20"; do timeout 120 python tools/msm_once.py $args 0 2>&1 | grep -E "curve|rror"; done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['value'], d['key_device_bytes'], d['roofline_int']['frac'])"
