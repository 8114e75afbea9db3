# Round-2 evidence on ONE B200: tests, bench lines (czk arm, reference arm, plonk), ncu launch list of the bench command, ncu
# --set full pages of the dominant kernels (exported to CSV on the box: the reports exceed the 64 MiB that travels back).
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' 2>&1 | tail -2 > gpurun_out/r2_smoke.log; cat gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; tail -c 300 gpurun_out/r2_bench_1gpu.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; cut -c1-400 gpurun_out/r2_bench_reference.json
timeout 300 python bench.py --workload plonk --log-n 18 --steps 5 --warmup 3 > gpurun_out/r2_bench_plonk_1gpu.json 2>/dev/null
timeout 300 python bench.py --workload plonk --log-n 18 --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_plonk_reference.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
for spec in "g1 1 21 22" "g2 2 20 22"; do
  set -- $spec
  timeout 400 ncu --set full --clock-control none -k regex:"k_bat_(round|finish)" -s $4 -c 22 -o /tmp/ncu/r2_acc_$1 -f python tools/msm_once.py $2 $3 0 > gpurun_out/r2_ncu_acc_$1.log 2>&1
  ncu -i /tmp/ncu/r2_acc_$1.ncu-rep --page raw --csv > gpurun_out/r2_acc_$1_raw.csv 2>/dev/null
done
timeout 300 ncu --set full --clock-control none -k regex:k_ntt_tile -s 12 -c 6 -o /tmp/ncu/r2_ntt -f python tools/ntt_once.py 21 > gpurun_out/r2_ncu_ntt.log 2>&1
ncu -i /tmp/ncu/r2_ntt.ncu-rep --page raw --csv > gpurun_out/r2_ntt_raw.csv 2>/dev/null
timeout 200 python tools/mb.py > gpurun_out/mb.log 2>&1
timeout 120 python tools/ntt_once.py 21 > gpurun_out/r2_ntt_once.log 2>&1
ls -la gpurun_out | head -40
