mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ntt_tile -s 12 -c 6 -o gpurun_out/r2_ntt_tile -f python tools/ntt_once.py 21 > gpurun_out/r2_ncu_ntt.log 2>&1
tail -3 gpurun_out/r2_ncu_ntt.log
ls -la gpurun_out/r2_ntt_tile.ncu-rep
