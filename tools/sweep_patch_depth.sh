#!/bin/bash
# Sweep load-pipeline depth x compiled occupancy of the inter patch kernels (ncu gpu__time_duration per launch of
# 2 x 4K pictures, us; columns luma-bi, luma-uni, chroma-bi, chroma-uni)
cd /root/repo
for d in ${DEPTHS:-2 4 6}; do
for cfg in "4 4 7 6" "3 3 5 5"; do
  set -- $cfg
  rm -f ffvvc_b200/csrc/build/inter_patch.o
  make -s -C ffvvc_b200/csrc EXTRA="-DPATCH_DEPTH=$d -DPATCH_MB_LU=$1 -DPATCH_MB_LB=$2 -DPATCH_MB_CU=$3 -DPATCH_MB_CB=$4" > /dev/null 2>&1
  echo -n "DEPTH=$d LU=$1 LB=$2 CU=$3 CB=$4 : "
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:inter_patch -s 4 -c 4 python tools/profile_recon.py 2 2 2>&1 | grep -E "gpu__time" | awk '{printf "%s ", $3} END {print ""}'
done
done
rm -f ffvvc_b200/csrc/build/inter_patch.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
