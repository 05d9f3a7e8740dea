#!/bin/bash
# Sweep the compiled occupancy (resident CTAs per SM) of the inter patch kernels on the GPU box.
cd /root/repo
for cfg in "4 4 5 5" "5 5 6 6" "6 5 7 6" "6 6 8 8" "8 6 8 8" "3 3 4 4"; do
  set -- $cfg
  rm -f ffvvc_b200/csrc/build/inter_patch.o
  make -s -C ffvvc_b200/csrc EXTRA="-DPATCH_MB_LU=$1 -DPATCH_MB_LB=$2 -DPATCH_MB_CU=$3 -DPATCH_MB_CB=$4" > /dev/null 2>&1
  echo "cfg LU=$1 LB=$2 CU=$3 CB=$4"
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:inter_patch -s 4 -c 4 python tools/profile_recon.py 2 2 2>&1 | grep -E "gpu__time" | awk '{printf "%s ", $3} END {print ""}'
done
rm -f ffvvc_b200/csrc/build/inter_patch.o; make -s -C ffvvc_b200/csrc > /dev/null 2>&1
