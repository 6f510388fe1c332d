#!/bin/bash
# experiment: does the SIZE of the kernel (not only the hot set) matter for the command path?  alt = the same kernel without the
# literal fast loops compiled in (7 K instead of 10 K instructions); Z population at 4096 streams, A/B on one box
cd /root/repo; mkdir -p gpurun_out; T=r2_v14
cp divans_b200/lib/libdivans_b200.so /tmp/new.so; cp divans_b200/lib/alt/libdivans_b200.so /tmp/alt.so
for round in 1 2; do for V in new alt; do
  cp /tmp/$V.so divans_b200/lib/libdivans_b200.so
  echo "== $V (round $round)" | tee -a gpurun_out/${T}_zsize.txt
  timeout 300 python tools/zprobe.py 4096 2>&1 | tail -1 | tee -a gpurun_out/${T}_zsize.txt
done; done
cp /tmp/new.so divans_b200/lib/libdivans_b200.so
timeout 900 python -m pytest tests/test_gpu_blend.py -m gpu -x -q 2>&1 | tail -3
