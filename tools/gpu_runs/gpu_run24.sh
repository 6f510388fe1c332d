#!/bin/bash
# the 8-lane kernel at the N > 1 workload: r2.10 (literal loop not unrolled) vs r2.8 (unrolled by 2), one box
cd /root/repo; mkdir -p gpurun_out; T=r2_v19
cp divans_b200/lib/libdivans_b200.so /tmp/new.so; cp divans_b200/lib/alt/libdivans_b200.so /tmp/alt.so
for round in 1 2 3; do for V in new alt; do
  cp /tmp/$V.so divans_b200/lib/libdivans_b200.so
  echo "== $V (round $round)" | tee -a gpurun_out/${T}_ab8.txt
  DIVANS_B200_LPS=0 timeout 300 python tools/perf_probe.py --l-only --decode-once 8192 2>&1 | cut -c1-120 | tee -a gpurun_out/${T}_ab8.txt
done; done
cp /tmp/new.so divans_b200/lib/libdivans_b200.so
