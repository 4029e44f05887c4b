#!/bin/bash
# run tools/profile_step.py against each kernel-variant library in gpclust_b200/_lib/var
for f in gpclust_b200/_lib/libgpclust_b200.so gpclust_b200/_lib/var/*.so; do
  echo "== $f"
  GPCLUST_B200_LIB=$PWD/$f timeout 200 python tools/profile_step.py ${1:-1000000} ${2:-30} 2>&1 | tail -1
done
