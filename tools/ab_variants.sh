#!/bin/bash
# A/B of msm_batched.cu build variants on ONE GPU box (box-to-box variance is ~10 %, so both arms must share a box).
#   here (build container):   tools/ab_variants.sh build bingcd=-DBAT_BINGCD=1 [name=flags ...]
#   on the box (via gpurun):  gpurun -- 'bash tools/ab_variants.sh run bingcd [name ...] | tee gpurun_out/ab.log'
# `run` times G1 2^21 / 2^20 and G2 2^20 accumulations for the shipped library and every named variant, then runs the MSM
# parity tests against each variant (CZK_B200_LIB).  variants/ is git-ignored but travels with the snapshot
# (~25 MB per variant): delete it when done.
set -e
cd "$(dirname "$0")/.."
mode=$1; shift
if [ "$mode" = build ]; then
  for spec in "$@"; do
    name=${spec%%=*}; flags=${spec#*=}
    tools/build_variant.sh "$name" msm_batched.cu $flags
  done
  rm -f variants/*.o
  exit 0
fi
time_one() {  # lib label
  for args in "1 21" "1 20" "2 20"; do
    echo -n "$2: "; CZK_B200_LIB=$1 python tools/msm_once.py $args 0 2>&1 | grep curve
  done
}
time_one collaborative-zksnark_b200/libczk_b200.so shipped
for name in "$@"; do time_one variants/$name.so "$name"; done
time_one collaborative-zksnark_b200/libczk_b200.so shipped-again
for name in "$@"; do
  echo "== parity tests with variants/$name.so"
  CZK_B200_LIB=variants/$name.so timeout 300 python -m pytest tests/test_gpu_msm.py -x -q -m gpu 2>&1 | tail -2
done
