# ncu --set full of the bucket-accumulation kernels of ONE 2^21-term G1 MSM (rounds 0.. + finish) and of one 2^20+1-term G2
# MSM.  The reports are too large to travel back (64 MiB cap), so the CSV pages are exported on the box and the .ncu-rep dropped.
mkdir -p gpurun_out /tmp/ncu
for spec in "g1 1 21" "g2 2 20"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_bat_(round|finish)" -s 24 -c 7 -o /tmp/ncu/r2_acc_$1 -f python tools/msm_once.py $2 $3 0 > gpurun_out/r2_ncu_acc_$1.log 2>&1
  tail -1 gpurun_out/r2_ncu_acc_$1.log
  ncu -i /tmp/ncu/r2_acc_$1.ncu-rep --page raw --csv > gpurun_out/r2_acc_$1_raw.csv 2>/dev/null
  for k in 0 1; do
    ncu -i /tmp/ncu/r2_acc_$1.ncu-rep --page source --csv --launch-skip $k --launch-count 1 > gpurun_out/r2_acc_$1_src$k.csv 2>/dev/null
  done
done
ls -la gpurun_out/ /tmp/ncu
