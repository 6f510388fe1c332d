#!/bin/bash
# the literal loop's schedule moves by +-2 % with unrelated changes of the kernel: variants on one box (L at 4096 streams, Z)
cd /root/repo; mkdir -p gpurun_out; T=r2_v18
cp divans_b200/lib/libdivans_b200.so /tmp/keep.so
for round in 1 2; do for V in r28 r29 r29_unroll1 r29_unroll4; do
  cp divans_b200/lib/var/$V.so divans_b200/lib/libdivans_b200.so
  echo "== $V (round $round)" | tee -a gpurun_out/${T}_variants.txt
  timeout 300 python tools/perf_probe.py --l-only --decode-once 4096 2>&1 | cut -c1-120 | tee -a gpurun_out/${T}_variants.txt
  if [ $round = 1 ]; then timeout 300 python tools/zprobe.py 4096 2>&1 | tail -1 | tee -a gpurun_out/${T}_variants.txt; fi
done; done
cp /tmp/keep.so divans_b200/lib/libdivans_b200.so
